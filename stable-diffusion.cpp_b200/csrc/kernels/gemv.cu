// gemv.cu -- MUL_MAT with a handful of activation rows (the timestep / label embedding MLPs and the per-ResBlock
// `Linear(SiLU(emb))` of the UNet: src/model/diffusion/unet.hpp:590-600, src/model/common/block.hpp:124-181).
//
// With N <= 4 rows a 128 x BN tensor-core tile is > 96 % padding and the launch is bound by streaming the weight matrix once, so
// this is an HBM-roofline kernel: one warp per output feature, 16-byte weight loads, the (optionally SiLU'd) activation rows staged
// once per block in shared memory ALREADY ROUNDED to the weight type -- the reference CPU path converts src1 to src0's vec_dot
// type before the dot product (ggml-cpu.c:1430-1513) -- f32 accumulation, bias / residual in the epilogue.
// Algorithmic bytes: M*K*2 (weights) + N*K*4 + N*M*4.
#include "../b200_ops.h"
#include "b200_launch.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

namespace {

constexpr int kWarps = 4;
constexpr int kMaxN = 4;

template <typename WT> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<__half>(float v) { return __half2float(__float2half_rn(v)); }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <typename WT> __device__ __forceinline__ void unpack8(const uint4& u, float* w);
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& u, float* w) {
    const __half2* h = (const __half2*)&u;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); w[2 * i] = f.x; w[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& u, float* w) {
    const __nv_bfloat162* h = (const __nv_bfloat162*)&u;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); w[2 * i] = f.x; w[2 * i + 1] = f.y; }
}

template <typename WT>
__global__ void __launch_bounds__(kWarps * 32) k_gemv(const WT* __restrict__ W, int64_t lda, const float* __restrict__ X, int64_t ldx, float* __restrict__ D,
                                                     int64_t ldd, int M, int K, int N, const float* __restrict__ bias, const float* __restrict__ residual,
                                                     int64_t ldr, int pre_act, int rows_per_warp) {
    extern __shared__ float xs[];    // [N][K], rounded to WT
    pdl_wait();
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a block owns kWarps * rows_per_warp consecutive output features; warp w takes rows m0 + w, m0 + w + kWarps, ...  The activation rows
    // are staged (read, SiLU, rounded) ONCE per block: with one row per warp that staging cost as much traffic and more instructions
    // than the four weight rows it served (Flux modulation: 12 KB of activations per 24 KB of weights).
    const int m0 = blockIdx.x * kWarps * rows_per_warp;
    const int K8 = K / 8;
    // the weight stream is what this kernel waits for: put the first 4 x 16 bytes per lane in flight BEFORE the activation rows are staged
    // (their global reads + SiLU + barrier used to sit in front of the first weight load), then keep 4 loads per lane ahead of the math
    int m = m0 + warp;
    const uint4* wrow = (const uint4*)(W + (int64_t)min(m, M - 1) * lda);
    uint4 wreg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = lane + 32 * j;
        wreg[j] = c < K8 ? wrow[c] : make_uint4(0u, 0u, 0u, 0u);
    }
    if ((((uintptr_t)X) & 15) == 0 && (ldx & 3) == 0) {
        // 16-byte loads, four per thread in flight: the scalar loop below is a chain of dependent-latency iterations (load, SiLU, store)
        // and cost more than streaming the block's weight rows (K % 8 == 0: a float4 never straddles two activation rows)
        const int K4 = K >> 2, total4 = N * K4;
        for (int i0 = threadIdx.x; i0 < total4; i0 += 4 * blockDim.x) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < total4) { const int n = i / K4, k4 = i - n * K4; q[u] = *(const float4*)(X + (int64_t)n * ldx + 4 * k4); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < total4) {
                    float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v = e[c];
                        if (pre_act == 1) v = v / (1.0f + expf(-v));     // SiLU, the unary kernel's own expression
                        e[c] = round_to<WT>(v);
                    }
                    *(float4*)(xs + 4 * i) = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
            const int n = i / K, k = i - n * K;
            float v = X[(int64_t)n * ldx + k];
            if (pre_act == 1) v = v / (1.0f + expf(-v));     // SiLU, the unary kernel's own expression
            xs[i] = round_to<WT>(v);
        }
    }
    __syncthreads();
    for (int r = 0; r < rows_per_warp; ++r, m += kWarps) {
        if (m >= M) return;
        // the NEXT row of this warp: its first chunks are issued before this row's reduction and store
        const int mn = m + kWarps;
        const uint4* wnext = (const uint4*)(W + (int64_t)min(mn, M - 1) * lda);
        const bool has_next = r + 1 < rows_per_warp && mn < M;
        float acc[kMaxN] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = lane; c0 < K8; c0 += 128) {
            uint4 cur[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) cur[j] = wreg[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + 128 + 32 * j;
                wreg[j] = c < K8 ? wrow[c] : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + 32 * j;
                if (c < K8) {
                    float w[8];
                    unpack8<WT>(cur[j], w);
#pragma unroll
                    for (int n = 0; n < kMaxN; ++n) {
                        if (n < N) {
                            const float4 x0 = *(const float4*)(xs + n * K + c * 8), x1 = *(const float4*)(xs + n * K + c * 8 + 4);
                            const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[n] = fmaf(w[i], x[i], acc[n]);
                        }
                    }
                }
            }
        }
        if (has_next) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * j;
                wreg[j] = c < K8 ? wnext[c] : make_uint4(0u, 0u, 0u, 0u);
            }
            wrow = wnext;
        }
#pragma unroll
        for (int n = 0; n < kMaxN; ++n) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
        }
        if (lane == 0) {
            const float b = bias ? bias[m] : 0.f;
            for (int n = 0; n < N; ++n) {
                float v = acc[n] + b;
                if (residual) v += residual[(int64_t)n * ldr + m];
                D[(int64_t)n * ldd + m] = v;
            }
        }
    }
}

}  // namespace

bool b200_gemv_supported(int wtype, int64_t M, int64_t N, int64_t K, const void* W, int64_t lda, const void* X) {
    if (wtype != GGML_TYPE_F16 && wtype != GGML_TYPE_BF16) return false;
    if (N < 1 || N > kMaxN || M < 1 || M > 0x7fffffff || K < 8 || K % 8) return false;
    if (N * K * 4 > 48 * 1024) return false;
    if (((uintptr_t)W & 15) || (lda * 2) % 16 || ((uintptr_t)X & 3)) return false;
    return true;
}

int b200_launch_gemv(cudaStream_t s, int wtype, const void* W, int64_t lda, const float* X, int64_t ldx, float* D, int64_t ldd, int64_t M, int64_t N,
                     int64_t K, const float* bias, const float* residual, int64_t ldr, int pre_act) {
    if (!b200_gemv_supported(wtype, M, N, K, W, lda, X)) return -1;
    // rows per warp: as many as keep >= 4 blocks per SM (the UNet's 320..1280-feature embeddings stay at one row per warp; the Flux / SD3
    // modulation Linears with 9216..18432 features take 4..8)
    static int rpw_env = -1;
    if (rpw_env < 0) { const char* e = getenv("GGML_B200_GEMV_RPW"); rpw_env = (e && *e) ? atoi(e) : 0; }
    int rpw = rpw_env > 0 ? rpw_env : (int)std::min<int64_t>(8, std::max<int64_t>(1, M / (kWarps * 148 * 4)));
    const dim3 grid((unsigned)((M + kWarps * rpw - 1) / (kWarps * rpw)));
    const size_t smem = (size_t)(N * K * 4);
    if (wtype == GGML_TYPE_F16)
        b200_launch(k_gemv<__half>, grid, dim3(kWarps * 32), smem, s, (const __half*)W, lda, X, ldx, D, ldd, (int)M, (int)K, (int)N, bias, residual, ldr, pre_act, rpw);
    else
        b200_launch(k_gemv<__nv_bfloat16>, grid, dim3(kWarps * 32), smem, s, (const __nv_bfloat16*)W, lda, X, ldx, D, ldd, (int)M, (int)K, (int)N, bias,
                    residual, ldr, pre_act, rpw);
    return 1;
}
