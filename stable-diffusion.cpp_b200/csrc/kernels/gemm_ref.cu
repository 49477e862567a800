// gemm_ref.cu -- CUDA-core tiled GEMM.  NOT the product path: it exists (a) as the in-backend numerical
// cross-check for the tcgen05 kernel (option "tc_gemm"=0) and (b) for operand shapes the TMA path cannot
// describe (rows not 16-byte aligned, K < 8).  D[n][m] = sum_k A[m][k] * B[n][k], f32 accumulate.
#include "../b200_ops.h"
#include "b200_launch.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace {

template <typename T> __device__ __forceinline__ float ld(const void* p, int64_t i);
template <> __device__ __forceinline__ float ld<float>(const void* p, int64_t i) { return ((const float*)p)[i]; }
template <> __device__ __forceinline__ float ld<__half>(const void* p, int64_t i) { return __half2float(((const __half*)p)[i]); }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const void* p, int64_t i) { return __bfloat162float(((const __nv_bfloat16*)p)[i]); }

constexpr int TM = 64, TN = 64, TK = 16;

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) k_gemm_ref(const char* __restrict__ A, int64_t lda, const char* __restrict__ B, int64_t ldb,
                                                  float* __restrict__ D, int64_t ldd, int64_t M, int64_t N, int64_t K) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float sa[TK][TM + 1];
    __shared__ float sb[TK][TN + 1];
    int64_t m0 = (int64_t)blockIdx.x * TM, n0 = (int64_t)blockIdx.y * TN;
    int tx = threadIdx.x % 16, ty = threadIdx.x / 16;   // thread computes rows m = tx*4..+3, cols n = ty*4..+3
    float acc[4][4] = {};
    for (int64_t k0 = 0; k0 < K; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            int kk = i % TK, mm = i / TK;
            int64_t m = m0 + mm, k = k0 + kk;
            sa[kk][mm] = (m < M && k < K) ? ld<TA>(A + m * lda, k) : 0.f;
        }
        for (int i = threadIdx.x; i < TN * TK; i += 256) {
            int kk = i % TK, nn = i / TK;
            int64_t n = n0 + nn, k = k0 + kk;
            sb[kk][nn] = (n < N && k < K) ? ld<TB>(B + n * ldb, k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sa[kk][tx * 4 + i]; b[i] = sb[kk][ty * 4 + i]; }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(a[i], b[j], acc[j][i]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int64_t n = n0 + ty * 4 + j;
        if (n >= N) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int64_t m = m0 + tx * 4 + i;
            if (m < M) D[n * ldd + m] = acc[j][i];
        }
    }
}

}  // namespace

int b200_launch_gemm_ref(cudaStream_t s, const void* A, int a_type, int64_t lda_bytes, const void* B, int b_type, int64_t ldb_bytes, float* D,
                         int64_t ldd, int64_t M, int64_t N, int64_t K) {
    if (M == 0 || N == 0) return 0;
    dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((N + TN - 1) / TN));
#define G(TA, TB) b200_launch(k_gemm_ref<TA, TB>, dim3(grid), dim3(256), 0, s, (const char*)A, lda_bytes, (const char*)B, ldb_bytes, D, ldd, M, N, K)
    if (a_type == GGML_TYPE_F32 && b_type == GGML_TYPE_F32) G(float, float);
    else if (a_type == GGML_TYPE_F16 && b_type == GGML_TYPE_F16) G(__half, __half);
    else if (a_type == GGML_TYPE_BF16 && b_type == GGML_TYPE_BF16) G(__nv_bfloat16, __nv_bfloat16);
    else if (a_type == GGML_TYPE_F16 && b_type == GGML_TYPE_F32) G(__half, float);
    else if (a_type == GGML_TYPE_BF16 && b_type == GGML_TYPE_F32) G(__nv_bfloat16, float);
    else if (a_type == GGML_TYPE_F32 && b_type == GGML_TYPE_F16) G(float, __half);
    else return -1;
#undef G
    return 1;
}
