// norm.cu -- reductions along rows / groups: GROUP_NORM, NORM, RMS_NORM, L2_NORM, SOFT_MAX.
//
// Oracle semantics (reference CPU backend):
//   group_norm  ggml/src/ggml-cpu/ops.cpp:4079-4152  two passes (mean, then centred variance), eps inside sqrt,
//               group g covers channels [g*cpg, min((g+1)*cpg, C)) with cpg = ceil(C / n_groups)
//   norm        ops.cpp (ggml_compute_forward_norm_f32): mean, centred variance, 1/sqrt(var + eps)
//   rms_norm    ops.cpp: 1/sqrt(mean(x^2) + eps);  l2_norm: 1/max(sqrt(sum x^2), eps)
//   soft_max    ops.cpp (ggml_compute_forward_soft_max_f32): x*scale + slope*mask, max-subtracted exp, / sum
// All are HBM-bound (one read + one write of the tensor is the algorithmic traffic); rows/groups are kept
// in registers or shared memory between the statistics pass and the normalise pass so HBM sees each
// element once on the way in and once on the way out.
#include "../b200_ops.h"
#include "b200_launch.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cfloat>

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide reductions; `red` is >= 32 floats of shared memory; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (w == 0) r = warp_sum(r);
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (threadIdx.x < nw) ? red[threadIdx.x] : -FLT_MAX;
    if (w == 0) r = warp_max(r);
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    return red[0];
}

// ------------------------------------------------------------------------------------------
// GROUP_NORM: one CTA per (group, batch).  The group is a contiguous run of `len` floats when the
// tensor is contiguous (W*H*cpg); it is staged in dynamic shared memory when it fits (<= 220 KB),
// so HBM is read once; larger groups are re-read from L2/HBM in the second and third pass.
// ------------------------------------------------------------------------------------------
template <bool IN_SMEM>
__global__ void __launch_bounds__(1024) k_group_norm(const float* __restrict__ x, float* __restrict__ y, int64_t inner /*W*H*/,
                                                     int C, int cpg, int n_groups, float eps, const float* __restrict__ gw,
                                                     const float* __restrict__ gb, int act) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float sbuf[];
    __shared__ float red[32];
    int g = blockIdx.x, n = blockIdx.y;
    int c0 = g * cpg, c1 = min(c0 + cpg, C);
    if (c0 >= c1) return;
    int64_t len = (int64_t)(c1 - c0) * inner;
    const float* xp = x + ((int64_t)n * C + c0) * inner;
    float* yp = y + ((int64_t)n * C + c0) * inner;

    float s = 0.f;
    if (((uintptr_t)xp % 16 == 0) && (len % 4 == 0)) {
        const float4* x4 = (const float4*)xp;
        for (int64_t i = threadIdx.x; i < len / 4; i += blockDim.x) {
            float4 v = x4[i];
            if (IN_SMEM) ((float4*)sbuf)[i] = v;
            s += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
            float v = xp[i];
            if (IN_SMEM) sbuf[i] = v;
            s += v;
        }
    }
    float mean = block_sum(s, red) / (float)len;
    float s2 = 0.f;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
        float v = (IN_SMEM ? sbuf[i] : xp[i]) - mean;
        s2 += v * v;
    }
    float var = block_sum(s2, red) / (float)len;
    float scale = 1.0f / sqrtf(var + eps);
    if (gw != nullptr || act != 0) {
        // fused epilogue of the reference's GroupNorm32 block: (x - mean) * rstd * w[c] + b[c], then SiLU
        // (ggml_extend.hpp:1502-1520 + block.hpp:142: GROUP_NORM -> MUL -> ADD -> SILU as four graph nodes)
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
            const int c = c0 + (int)(i / inner);
            float v = ((IN_SMEM ? sbuf[i] : xp[i]) - mean) * scale;
            if (gw) v = v * gw[c] + (gb ? gb[c] : 0.f);
            if (act == 1) v = v / (1.0f + expf(-v));
            yp[i] = v;
        }
        return;
    }
    if (((uintptr_t)yp % 16 == 0) && (len % 4 == 0) && ((uintptr_t)xp % 16 == 0)) {
        for (int64_t i = threadIdx.x; i < len / 4; i += blockDim.x) {
            float4 v = IN_SMEM ? ((const float4*)sbuf)[i] : ((const float4*)xp)[i];
            ((float4*)yp)[i] = make_float4((v.x - mean) * scale, (v.y - mean) * scale, (v.z - mean) * scale, (v.w - mean) * scale);
        }
    } else {
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) yp[i] = ((IN_SMEM ? sbuf[i] : xp[i]) - mean) * scale;
    }
}

// ------------------------------------------------------------------------------------------
// row norms: one CTA per row (ne0 elements, arbitrary row placement, unit stride inside the row)
// ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void k_row_norm(b200_td a, b200_td d, float eps, const float* __restrict__ rw, const float* __restrict__ rb, void* out16, int out16_bf16,
                           int modulate) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[32];
    int64_t row = blockIdx.x;
    int64_t i1 = row % a.ne[1], r = row / a.ne[1];
    int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
    const float* x = (const float*)((const char*)a.data + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
    float* y = (float*)((char*)d.data + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]);
    int64_t n = a.ne[0];
    // rows up to 8 elements per thread * blockDim are cached in registers
    constexpr int MAXR = 8;
    float v[MAXR];
    bool cached = n <= (int64_t)MAXR * blockDim.x;
    float s = 0.f;
    if (cached) {
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
            int64_t i = threadIdx.x + (int64_t)k * blockDim.x;
            v[k] = i < n ? x[i] : 0.f;
            s += (KIND == B200_NORM_LAYER) ? v[k] : v[k] * v[k];
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { float t = x[i]; s += (KIND == B200_NORM_LAYER) ? t : t * t; }
    }
    s = block_sum(s, red);
    float mean = 0.f, scale;
    if (KIND == B200_NORM_LAYER) {
        mean = s / (float)n;
        float s2 = 0.f;
        if (cached) {
#pragma unroll
            for (int k = 0; k < MAXR; ++k) {
                int64_t i = threadIdx.x + (int64_t)k * blockDim.x;
                if (i < n) { float t = v[k] - mean; s2 += t * t; }
            }
        } else {
            for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { float t = x[i] - mean; s2 += t * t; }
        }
        s2 = block_sum(s2, red);
        scale = 1.0f / sqrtf(s2 / (float)n + eps);
    } else if (KIND == B200_NORM_RMS) {
        scale = 1.0f / sqrtf(s / (float)n + eps);
    } else {
        scale = 1.0f / fmaxf(sqrtf(s), eps);
    }
    if (cached) {
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
            int64_t i = threadIdx.x + (int64_t)k * blockDim.x;
            if (i < n) {
                float o = (v[k] - mean) * scale;
                if (rw) o = modulate ? __fadd_rn(__fadd_rn(o, __fmul_rn(o, rw[i])), rb[i]) : o * rw[i] + (rb ? rb[i] : 0.f);
                y[i] = o;
                if (out16) {
                    if (out16_bf16) ((__nv_bfloat16*)out16)[row * n + i] = __float2bfloat16_rn(o);
                    else ((__half*)out16)[row * n + i] = __float2half_rn(o);
                }
            }
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            float o = (x[i] - mean) * scale;
            if (rw) o = modulate ? __fadd_rn(__fadd_rn(o, __fmul_rn(o, rw[i])), rb[i]) : o * rw[i] + (rb ? rb[i] : 0.f);
            y[i] = o;
            if (out16) {
                if (out16_bf16) ((__nv_bfloat16*)out16)[row * n + i] = __float2bfloat16_rn(o);
                else ((__half*)out16)[row * n + i] = __float2half_rn(o);
            }
        }
    }
}

// Vectorised row norm: rows of up to 4096 floats live in registers as float4 (one 16-byte load per lane and slot, coalesced 4 KB per
// CTA-wide instruction), f32 result stored as float4 and the optional 16-bit operand copy as 8 bytes per lane.  The scalar kernel above
// re-read rows longer than 2048 floats three times with 4-byte accesses: 0.5-1.4 TB/s on the DiT blocks' [3072, 4352] LayerNorms.
template <int KIND>
__global__ void __launch_bounds__(256) k_row_norm_vec(b200_td a, b200_td d, float eps, const float* __restrict__ rw, const float* __restrict__ rb, void* out16,
                                                      int out16_bf16, int modulate) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[32];
    const int64_t row = blockIdx.x;
    const int64_t i1 = row % a.ne[1], r = row / a.ne[1];
    const int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
    const float4* x = (const float4*)((const char*)a.data + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
    float4* y = (float4*)((char*)d.data + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]);
    const int n4 = (int)(a.ne[0] >> 2);
    const float n = (float)a.ne[0];
    float4 v[4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = threadIdx.x + k * 256;
        v[k] = i < n4 ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (KIND == B200_NORM_LAYER) ? (v[k].x + v[k].y) + (v[k].z + v[k].w) : (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
    }
    s = block_sum(s, red);
    float mean = 0.f, scale;
    if (KIND == B200_NORM_LAYER) {
        mean = s / n;
        float s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = threadIdx.x + k * 256;
            if (i < n4) {
                const float t0 = v[k].x - mean, t1 = v[k].y - mean, t2 = v[k].z - mean, t3 = v[k].w - mean;
                s2 += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
            }
        }
        s2 = block_sum(s2, red);
        scale = 1.0f / sqrtf(s2 / n + eps);
    } else if (KIND == B200_NORM_RMS) {
        scale = 1.0f / sqrtf(s / n + eps);
    } else {
        scale = 1.0f / fmaxf(sqrtf(s), eps);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < n4) {
            float o[4] = {(v[k].x - mean) * scale, (v[k].y - mean) * scale, (v[k].z - mean) * scale, (v[k].w - mean) * scale};
            if (rw) {
                const float4 w4 = ((const float4*)rw)[i];
                const float4 b4 = rb ? ((const float4*)rb)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float ww[4] = {w4.x, w4.y, w4.z, w4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = modulate ? __fadd_rn(__fadd_rn(o[e], __fmul_rn(o[e], ww[e])), bb[e]) : o[e] * ww[e] + bb[e];
            }
            y[i] = make_float4(o[0], o[1], o[2], o[3]);
            if (out16) {
                uint2 h;
                if (out16_bf16) {
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
                    h.x = *(const uint32_t*)&p0; h.y = *(const uint32_t*)&p1;
                } else {
                    const __half2 p0 = __floats2half2_rn(o[0], o[1]), p1 = __floats2half2_rn(o[2], o[3]);
                    h.x = *(const uint32_t*)&p0; h.y = *(const uint32_t*)&p1;
                }
                ((uint2*)out16)[row * n4 + i] = h;
            }
        }
    }
}

// Short rows (< 512 floats: the C = 320 LayerNorms of the 64 x 64 UNet level, 8192 rows per batched forward): ONE WARP per row, eight rows
// per CTA, the row in registers as float4, shuffle reductions, no block barrier.  The CTA-per-row kernel above spent 19 us on 26 MB
// (8192 CTAs of 128 threads, two block reductions each: 1.4 TB/s).
template <int KIND>
__global__ void __launch_bounds__(256) k_row_norm_warp(b200_td a, b200_td d, float eps, const float* __restrict__ rw, const float* __restrict__ rb, void* out16,
                                                       int out16_bf16, int modulate, int64_t nrows) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int64_t i1 = row % a.ne[1], r = row / a.ne[1];
    const int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
    const float4* x = (const float4*)((const char*)a.data + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
    float4* y = (float4*)((char*)d.data + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]);
    const int n4 = (int)(a.ne[0] >> 2);
    const float n = (float)a.ne[0];
    float4 v[4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = lane + k * 32;
        v[k] = i < n4 ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (KIND == B200_NORM_LAYER) ? (v[k].x + v[k].y) + (v[k].z + v[k].w) : (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
    }
    s = warp_sum(s);
    float mean = 0.f, scale;
    if (KIND == B200_NORM_LAYER) {
        mean = s / n;
        float s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = lane + k * 32;
            if (i < n4) {
                const float t0 = v[k].x - mean, t1 = v[k].y - mean, t2 = v[k].z - mean, t3 = v[k].w - mean;
                s2 += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
            }
        }
        s2 = warp_sum(s2);
        scale = 1.0f / sqrtf(s2 / n + eps);
    } else if (KIND == B200_NORM_RMS) {
        scale = 1.0f / sqrtf(s / n + eps);
    } else {
        scale = 1.0f / fmaxf(sqrtf(s), eps);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = lane + k * 32;
        if (i < n4) {
            float o[4] = {(v[k].x - mean) * scale, (v[k].y - mean) * scale, (v[k].z - mean) * scale, (v[k].w - mean) * scale};
            if (rw) {
                const float4 w4 = ((const float4*)rw)[i];
                const float4 b4 = rb ? ((const float4*)rb)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float ww[4] = {w4.x, w4.y, w4.z, w4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = modulate ? __fadd_rn(__fadd_rn(o[e], __fmul_rn(o[e], ww[e])), bb[e]) : o[e] * ww[e] + bb[e];
            }
            y[i] = make_float4(o[0], o[1], o[2], o[3]);
            if (out16) {
                uint2 h;
                if (out16_bf16) {
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
                    h.x = *(const uint32_t*)&p0; h.y = *(const uint32_t*)&p1;
                } else {
                    const __half2 p0 = __floats2half2_rn(o[0], o[1]), p1 = __floats2half2_rn(o[2], o[3]);
                    h.x = *(const uint32_t*)&p0; h.y = *(const uint32_t*)&p1;
                }
                ((uint2*)out16)[row * n4 + i] = h;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// SOFT_MAX: one CTA per row; the scaled+masked row lives in shared memory (ne0 floats).
// ------------------------------------------------------------------------------------------
template <typename TM>
__global__ void k_soft_max(b200_td a, b200_td m, b200_td d, bool has_mask, float scale, float max_bias, float m0, float m1, uint32_t n_head_log2) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float srow[];
    __shared__ float red[32];
    int64_t row = blockIdx.x;
    int64_t i1 = row % a.ne[1], r = row / a.ne[1];
    int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
    const float* x = (const float*)((const char*)a.data + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
    float* y = (float*)((char*)d.data + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]);
    const TM* mp = has_mask ? (const TM*)((const char*)m.data + i1 * m.nb[1] + (i2 % m.ne[2]) * m.nb[2] + (i3 % m.ne[3]) * m.nb[3]) : nullptr;
    float slope = 1.0f;
    if (max_bias > 0.0f) {
        uint32_t h = (uint32_t)i2;
        slope = h < n_head_log2 ? powf(m0, (float)(h + 1)) : powf(m1, (float)(2 * (h - n_head_log2) + 1));
    }
    int64_t n = a.ne[0];
    float mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float v = x[i] * scale;
        if (has_mask) v += slope * (float)mp[i];
        srow[i] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float e = expf(srow[i] - mx);
        srow[i] = e;
        s += e;
    }
    s = block_sum(s, red);
    float inv = 1.0f / s;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) y[i] = srow[i] * inv;
}

}  // namespace

int b200_launch_group_norm(cudaStream_t s, const b200_td& src, const b200_td& dst, int n_groups, float eps, const float* w, const float* b, int act) {
    int64_t inner = src.ne[0] * src.ne[1];
    int C = (int)src.ne[2];
    int N = (int)src.ne[3];
    if (inner * C * N == 0) return 0;
    int cpg = (C + n_groups - 1) / n_groups;
    size_t bytes = (size_t)cpg * inner * sizeof(float);
    dim3 grid(n_groups, N);
    int threads = (int)std::min<int64_t>(1024, std::max<int64_t>(128, ((cpg * inner / 4) + 31) / 32 * 32));
    if (bytes <= 200 * 1024) {
        static bool attr_set[B200_MAX_DEVICES] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attr_set[dev]) {
            cudaFuncSetAttribute(k_group_norm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_set[dev] = true;
        }
        b200_launch(k_group_norm<true>, dim3(grid), dim3(threads), bytes, s, (const float*)src.data, (float*)dst.data, inner, C, cpg, n_groups, eps, w, b, act);
    } else {
        b200_launch(k_group_norm<false>, dim3(grid), dim3(1024), 0, s, (const float*)src.data, (float*)dst.data, inner, C, cpg, n_groups, eps, w, b, act);
    }
    return 1;
}

int b200_launch_norm(cudaStream_t s, int kind, const b200_td& src, const b200_td& dst, float eps, const float* w, const float* b, void* out16,
                     int out16_type, int modulate) {
    if (modulate && (!w || !b)) return -1;
    const int bf = out16_type == GGML_TYPE_BF16 ? 1 : 0;
    int64_t nrows = src.ne[1] * src.ne[2] * src.ne[3];
    if (nrows == 0 || src.ne[0] == 0) return 0;
    int threads = src.ne[0] >= 4096 ? 1024 : (src.ne[0] >= 1024 ? 256 : 128);
    if (nrows > 0x7fffffff) return -1;
    // float4 path: rows of 512 .. 4096 floats, 16-byte aligned rows / affine vectors / 8-byte aligned 16-bit copy
    const bool vec = src.ne[0] % 4 == 0 && src.ne[0] >= 512 && src.ne[0] <= 4096 && src.nb[0] == 4 && dst.nb[0] == 4 && !((uintptr_t)src.data & 15) && !((uintptr_t)dst.data & 15) &&
                     !(src.nb[1] & 15) && !(src.nb[2] & 15) && !(src.nb[3] & 15) && !(dst.nb[1] & 15) && !(dst.nb[2] & 15) && !(dst.nb[3] & 15) &&
                     !((uintptr_t)w & 15) && !((uintptr_t)b & 15) && !((uintptr_t)out16 & 7);
    // short rows: a warp per row (same alignment conditions as the float4 path)
    static int warp_rows = -1;
    if (warp_rows < 0) { const char* e = getenv("GGML_B200_NORM_WARP"); warp_rows = (e && *e) ? atoi(e) : 1; }
    const bool vec_small = warp_rows && src.ne[0] % 4 == 0 && src.ne[0] >= 64 && src.ne[0] < 512 && nrows >= 64 && src.nb[0] == 4 && dst.nb[0] == 4 && !((uintptr_t)src.data & 15) &&
                           !((uintptr_t)dst.data & 15) && !(src.nb[1] & 15) && !(src.nb[2] & 15) && !(src.nb[3] & 15) && !(dst.nb[1] & 15) && !(dst.nb[2] & 15) && !(dst.nb[3] & 15) &&
                           !((uintptr_t)w & 15) && !((uintptr_t)b & 15) && !((uintptr_t)out16 & 7);
    if (vec_small) {
        const unsigned grid = (unsigned)((nrows + 7) / 8);
        switch (kind) {
            case B200_NORM_LAYER: b200_launch(k_row_norm_warp<B200_NORM_LAYER>, dim3(grid), dim3(256), 0, s, src, dst, eps, w, b, out16, bf, modulate, nrows); break;
            case B200_NORM_RMS: b200_launch(k_row_norm_warp<B200_NORM_RMS>, dim3(grid), dim3(256), 0, s, src, dst, eps, w, b, out16, bf, modulate, nrows); break;
            default: b200_launch(k_row_norm_warp<B200_NORM_L2>, dim3(grid), dim3(256), 0, s, src, dst, eps, w, b, out16, bf, modulate, nrows); break;
        }
        return 1;
    }
    if (vec) {
        switch (kind) {
            case B200_NORM_LAYER: b200_launch(k_row_norm_vec<B200_NORM_LAYER>, dim3((unsigned)nrows), dim3(256), 0, s, src, dst, eps, w, b, out16, bf, modulate); break;
            case B200_NORM_RMS: b200_launch(k_row_norm_vec<B200_NORM_RMS>, dim3((unsigned)nrows), dim3(256), 0, s, src, dst, eps, w, b, out16, bf, modulate); break;
            default: b200_launch(k_row_norm_vec<B200_NORM_L2>, dim3((unsigned)nrows), dim3(256), 0, s, src, dst, eps, w, b, out16, bf, modulate); break;
        }
        return 1;
    }
    switch (kind) {
        case B200_NORM_LAYER: b200_launch(k_row_norm<B200_NORM_LAYER>, dim3((unsigned)nrows), dim3(threads), 0, s, src, dst, eps, w, b, out16, bf, modulate); break;
        case B200_NORM_RMS: b200_launch(k_row_norm<B200_NORM_RMS>, dim3((unsigned)nrows), dim3(threads), 0, s, src, dst, eps, w, b, out16, bf, modulate); break;
        default: b200_launch(k_row_norm<B200_NORM_L2>, dim3((unsigned)nrows), dim3(threads), 0, s, src, dst, eps, w, b, out16, bf, modulate); break;
    }
    return 1;
}

int b200_launch_soft_max(cudaStream_t s, const b200_td& src, const b200_td* mask, const b200_td& dst, float scale, float max_bias) {
    int64_t nrows = src.ne[1] * src.ne[2] * src.ne[3];
    if (nrows == 0 || src.ne[0] == 0) return 0;
    size_t bytes = (size_t)src.ne[0] * sizeof(float);
    if (bytes > 200 * 1024 || nrows > 0x7fffffff) return -1;
    uint32_t n_head = (uint32_t)src.ne[2];
    uint32_t n_head_log2 = 1u << (uint32_t)floorf(log2f((float)n_head));
    float m0 = powf(2.0f, -(max_bias) / n_head_log2), m1 = powf(2.0f, -(max_bias / 2.0f) / n_head_log2);
    int threads = src.ne[0] >= 2048 ? 512 : (src.ne[0] >= 256 ? 256 : 64);
    b200_td m = mask ? *mask : src;
    bool f16mask = mask && mask->type == GGML_TYPE_F16;
    static bool attr_set[B200_MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(k_soft_max<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_soft_max<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set[dev] = true;
    }
    if (f16mask) b200_launch(k_soft_max<__half>, dim3((unsigned)nrows), dim3(threads), bytes, s, src, m, dst, mask != nullptr, scale, max_bias, m0, m1, n_head_log2);
    else b200_launch(k_soft_max<float>, dim3((unsigned)nrows), dim3(threads), bytes, s, src, m, dst, mask != nullptr, scale, max_bias, m0, m1, n_head_log2);
    return 1;
}
