// conv_prep.cu -- producers of the operands the implicit-GEMM convolution (gemm_tc.cu, conv mode) reads through TMA:
//
//   k_gn_stats        per (image, group) mean / rstd of a GroupNorm input (oracle arithmetic: ops.cpp:4079-4152)
//   k_to_nhwc_f16     NCHW f32 activation -> NHWC f16 "shadow" image with the whole ResBlock prologue folded in:
//                     (x - mean_g) * rstd_g * w_c + b_c, SiLU, and optional nearest x2 upsampling
//                     (reference graph: GROUP_NORM, MUL, ADD, SILU[, UPSCALE], IM2COL -- ggml_extend.hpp:1502-1520,
//                     block.hpp:58-65,142).  Rounding to f16 happens exactly where the oracle rounds (im2col output is F16).
//   k_pack_conv_weight  [KW,KH,IC,OC] f16 (ggml) -> [OC][KH][KW][IC] f16 so that K runs (tap, channel) like the NHWC image
//
// All three are HBM/L2 bound: the activation is read once (twice from L2 when statistics are needed) and written once
// at half width; the 0.7 GB/forward of materialised im2col columns of the unfused graph disappear.
#include "../b200_ops.h"

#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
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

__global__ void __launch_bounds__(1024) k_gn_stats(const float* __restrict__ x, float2* __restrict__ stats, int64_t inner, int C, int cpg, int G,
                                                   float eps) {
    __shared__ float red[32];
    const int g = blockIdx.x, n = blockIdx.y;
    const int c0 = g * cpg, c1 = min(c0 + cpg, C);
    if (c0 >= c1) return;
    const int64_t len = (int64_t)(c1 - c0) * inner;
    const float* xp = x + ((int64_t)n * C + c0) * inner;
    float s = 0.f;
    if (((uintptr_t)xp % 16 == 0) && (len % 4 == 0)) {
        const float4* x4 = (const float4*)xp;
        for (int64_t i = threadIdx.x; i < len / 4; i += blockDim.x) { float4 v = x4[i]; s += (v.x + v.y) + (v.z + v.w); }
    } else {
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) s += xp[i];
    }
    const float mean = block_sum(s, red) / (float)len;
    float s2 = 0.f;
    if (((uintptr_t)xp % 16 == 0) && (len % 4 == 0)) {
        const float4* x4 = (const float4*)xp;
        for (int64_t i = threadIdx.x; i < len / 4; i += blockDim.x) {
            float4 v = x4[i];
            float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
            s2 += (a * a + b * b) + (c * c + d * d);
        }
    } else {
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) { float v = xp[i] - mean; s2 += v * v; }
    }
    const float var = block_sum(s2, red) / (float)len;
    if (threadIdx.x == 0) stats[(int64_t)n * G + g] = make_float2(mean, 1.0f / sqrtf(var + eps));
}

// grid (ceil(OH*OW / 32), C / 64, N), block (32, 8)
__global__ void __launch_bounds__(256) k_to_nhwc_f16(const float* __restrict__ x, __half* __restrict__ out, const float2* __restrict__ stats,
                                                     const float* __restrict__ gw, const float* __restrict__ gb, int C, int H, int W, int OH,
                                                     int OW, int up, int cpg, int G, int act) {
    __shared__ float tile[64][33];
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * 64;
    const int p0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t OHW = (int64_t)OH * OW;
    // ---- load: tx runs over 32 consecutive output pixels (coalesced along W), ty over channels
    const int64_t pix = (int64_t)p0 + tx;
    int64_t src_off = 0;
    const bool pvalid = pix < OHW;
    if (pvalid) {
        const int oy = (int)(pix / OW), ox = (int)(pix - (int64_t)oy * OW);
        src_off = (int64_t)(oy / up) * W + (ox / up);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = c0 + ty + k * 8;
        float v = 0.f;
        if (pvalid && c < C) {
            v = x[((int64_t)n * C + c) * H * W + src_off];
            if (stats) {
                const float2 st = stats[(int64_t)n * G + c / cpg];
                v = (v - st.x) * st.y;
            }
            if (gw) v = v * gw[c] + (gb ? gb[c] : 0.f);
            if (act == 1) v = v / (1.0f + expf(-v));
        }
        tile[ty + k * 8][tx] = v;
    }
    __syncthreads();
    // ---- store: tx runs over 32 channel pairs (64 channels = 128 contiguous bytes per pixel), ty over pixels
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int pl = ty + k * 8;
        const int64_t p = (int64_t)p0 + pl;
        if (p < OHW && c0 + 2 * tx + 1 < C + 1) {
            __half2 hv = __floats2half2_rn(tile[2 * tx][pl], tile[2 * tx + 1][pl]);
            *(__half2*)(out + ((int64_t)n * OHW + p) * C + c0 + 2 * tx) = hv;
        }
    }
}

__global__ void k_pack_conv_weight(const __half* __restrict__ w, __half* __restrict__ out, int KW, int KH, int IC, int64_t total) {
    // out[((oc * KH + kh) * KW + kw) * IC + ic] = w[((oc * IC + ic) * KH + kh) * KW + kw]
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int ic = (int)(r % IC); r /= IC;
        const int kw = (int)(r % KW); r /= KW;
        const int kh = (int)(r % KH);
        const int64_t oc = r / KH;
        out[i] = w[((oc * IC + ic) * KH + kh) * KW + kw];
    }
}

}  // namespace

int b200_launch_gn_stats(cudaStream_t s, const float* x, float* stats, int64_t N, int64_t C, int64_t inner, int n_groups, float eps) {
    const int cpg = (int)((C + n_groups - 1) / n_groups);
    dim3 grid((unsigned)n_groups, (unsigned)N);
    k_gn_stats<<<grid, 1024, 0, s>>>(x, (float2*)stats, inner, (int)C, cpg, n_groups, eps);
    return 1;
}

int b200_launch_to_nhwc_f16(cudaStream_t s, const float* x, void* out, int64_t N, int64_t C, int64_t H, int64_t W, int up, const float* stats,
                            int n_groups, const float* gw, const float* gb, int act) {
    if (C % 64 != 0) return -1;
    const int64_t OH = H * up, OW = W * up;
    const int cpg = n_groups > 0 ? (int)((C + n_groups - 1) / n_groups) : 1;
    dim3 grid((unsigned)((OH * OW + 31) / 32), (unsigned)(C / 64), (unsigned)N);
    if (grid.y > 65535 || N > 65535) return -1;
    k_to_nhwc_f16<<<grid, dim3(32, 8), 0, s>>>(x, (__half*)out, (const float2*)stats, gw, gb, (int)C, (int)H, (int)W, (int)OH, (int)OW, up, cpg, n_groups,
                                               act);
    return 1;
}

int b200_launch_pack_conv_weight(cudaStream_t s, const void* w, void* out, int KW, int KH, int64_t IC, int64_t OC) {
    const int64_t total = (int64_t)KW * KH * IC * OC;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_pack_conv_weight<<<(unsigned)blocks, 256, 0, s>>>((const __half*)w, (__half*)out, KW, KH, (int)IC, total);
    return 1;
}
