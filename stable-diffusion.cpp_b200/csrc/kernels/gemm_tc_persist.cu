// gemm_tc_persist.cu -- EXPERIMENTAL persistent variant of the tcgen05 GEMM (round-2 work in progress).
//
// NOT on the default path: reachable only with GGML_B200_PERSISTENT=1 / option "persistent_gemm", and it has not run on hardware yet
// (written after the round's GPU budget was spent).  It lives in its own translation unit so that the validated kernels of
// gemm_tc.cu keep their exact machine code (their epilogue turned out to be sensitive to one extra instruction,
// profiles/r01_summary.md).
//
// Idea (profiles/r01_gemm_notes.md): for a 20-k-block problem the epilogue costs as many warp samples as the main loop.  One CTA
// per SM walks tiles t = blockIdx.x, += gridDim.x with TWO accumulators in TMEM; the TMA producer and the MMA issuer run ahead into
// the next tile while the four epilogue warps drain the previous accumulator.
#include "../b200_ops.h"
#include "b200_launch.cuh"
#include "sm100_ptx.cuh"

#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>

using namespace sm100;

namespace {

constexpr int BM = 128;
constexpr int BK_BYTES = 128;
constexpr int A_STAGE_BYTES = BM * BK_BYTES;

struct PGemmParams {
    float* D;
    int64_t ldd, d_batch_stride;
    int64_t M, N;
    int num_k_blocks;
    int ne12;
    int r2, r3;
    const float* bias;
    int bias_mode;     // 0 none, 1 per m, 2 per n
    const float* residual;
    int64_t ldr, r_batch_stride;
};

template <int BN> struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK_BYTES;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN, int FMT>
__global__ void __launch_bounds__(192, 1) k_gemm_tc_persistent(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               const PGemmParams p, const int tiles_m, const int tiles_n, const int total_tiles) {
    using C = Cfg<BN>;
    constexpr int BK = FMT == 2 ? 32 : 64;
    constexpr int UMMA_K = FMT == 2 ? 8 : 16;
    constexpr int ACC_COLS = 2 * C::TMEM_COLS;          // two accumulators: 128 / 256 / 512 TMEM columns
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[C::STAGES], empty_bar[C::STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.num_k_blocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);     // one elected lane of each of the four epilogue warps
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, ACC_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();
    pdl_launch_dependents();

    // tile -> (m tile, n tile, batch): m fastest, so CTAs that run concurrently share the B (weight / filter) tile in L2
    auto decode = [&](int t, int& m0, int& n0, int& batch) {
        const int mt = t % tiles_m;
        const int r = t / tiles_m;
        const int nt = r % tiles_n;
        batch = r / tiles_n;
        m0 = mt * BM;
        n0 = nt * BN;
    };

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                int m0, n0, batch;
                decode(t, m0, n0, batch);
                const int i2 = batch % p.ne12, i3 = batch / p.ne12;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], C::STAGE_BYTES);
                    uint8_t* sa = smem + s * C::STAGE_BYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    const int k = kb * BK;
                    // (conv mode is not wired into this variant yet)
                    tma_load_4d(sa, &tmA, &full_bar[s], k, m0, i2 / p.r2, i3 / p.r3);
                    tma_load_4d(sb, &tmB, &full_bar[s], k, n0, i2, i3);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc(FMT, BM, BN);
        uint32_t it = 0, j = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
            const uint32_t buf = j & 1, aph = (j >> 1) & 1;
            mbar_wait(&acc_empty[buf], aph ^ 1);            // the epilogue has drained this accumulator (first two tiles: free)
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * C::TMEM_COLS;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % C::STAGES;
                const uint32_t ph = (it / C::STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
                    const uint64_t da = make_smem_desc_sw128(sa);
                    const uint64_t db = make_smem_desc_sw128(sa + A_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (FMT == 2) mma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        else mma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    mma_commit(&empty_bar[s]);
                    if (kb == nkb - 1) mma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;
        const int ml = q * 32 + lane;
        uint32_t j = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
            int m0, n0, batch;
            decode(t, m0, n0, batch);
            const uint32_t buf = j & 1, aph = (j >> 1) & 1;
            const int64_t m = (int64_t)m0 + ml;
            const bool mvalid = m < p.M;
            const float bias_m = (p.bias_mode == 1 && mvalid) ? p.bias[m] : 0.f;
            const int ncols = (int)min((int64_t)BN, p.N - n0);
            float* dptr = p.D + (int64_t)batch * p.d_batch_stride + (int64_t)n0 * p.ldd + m;
            const float* rptr = p.residual ? p.residual + (int64_t)batch * p.r_batch_stride + (int64_t)n0 * p.ldr + m : nullptr;
            const float* bias_n = p.bias_mode == 2 ? p.bias + n0 : nullptr;
            mbar_wait(&acc_full[buf], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * C::TMEM_COLS;
#pragma unroll 1
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                const float bn = (bias_n && c0 + lane < ncols) ? bias_n[c0 + lane] : 0.f;
                tmem_ld_wait();
                if (rptr == nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = __uint_as_float(r[i]) + bias_m + __shfl_sync(0xffffffffu, bn, i);
                        if (mvalid && c0 + i < ncols) dptr[(int64_t)(c0 + i) * p.ldd] = v;
                    }
                } else {
                    float rr[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) rr[i] = (mvalid && c0 + i < ncols) ? rptr[(int64_t)(c0 + i) * p.ldr] : 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = __uint_as_float(r[i]) + bias_m + __shfl_sync(0xffffffffu, bn, i) + rr[i];
                        if (mvalid && c0 + i < ncols) dptr[(int64_t)(c0 + i) * p.ldd] = v;
                    }
                }
            }
            // every tcgen05.ld of this buffer has completed (tmem_ld_wait above): hand the accumulator back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ACC_COLS);
    }
}


bool encode_operand(CUtensorMap* out, const void* ptr, int type, int64_t K, int64_t rows, int64_t ld_elems, int64_t b2, int64_t b2_stride, uint32_t box_rows) {
    auto enc = b200_get_tensormap_encoder();
    if (!enc) return false;
    const int64_t es = type == GGML_TYPE_F32 ? 4 : 2;
    CUtensorMapDataType dt = type == GGML_TYPE_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                   : (type == GGML_TYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)b2, 1};
    cuuint64_t strides[3] = {(cuuint64_t)(ld_elems * es), (cuuint64_t)(b2_stride * es), 0};
    if (b2 == 1 || strides[1] == 0) strides[1] = strides[0] * dims[1];
    strides[2] = strides[1] * dims[2];
    cuuint32_t box[4] = {(cuuint32_t)(BK_BYTES / es), box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(out, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int FMT>
cudaError_t launch_persistent(cudaStream_t s, int sms, const CUtensorMap& ta, const CUtensorMap& tb, const PGemmParams& kp, int tiles_m, int tiles_n,
                              int total_tiles) {
    using C = Cfg<BN>;
    static bool configured[B200_MAX_DEVICES] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (!configured[d]) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc_persistent<BN, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[d] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min(total_tiles, sms));
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    cfg.numAttrs = b200_launch_attrs(attr, 1);
    cfg.attrs = attr;
    return cudaLaunchKernelEx(&cfg, k_gemm_tc_persistent<BN, FMT>, ta, tb, kp, tiles_m, tiles_n, total_tiles);
}

}  // namespace

// returns 1 when launched, -1 when the problem is outside this variant's envelope (caller uses b200_launch_gemm_tc)
int b200_launch_gemm_tc_persistent(cudaStream_t s, const b200_device_info& dev, const b200_gemm_args& g) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0 || g.K <= 0 || g.act != 0) return -1;
    if (g.type != GGML_TYPE_F32 && g.type != GGML_TYPE_F16 && g.type != GGML_TYPE_BF16) return -1;
    const int64_t es = g.type == GGML_TYPE_F32 ? 4 : 2;
    if (((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15) || (g.lda * es) % 16 || (g.ldb * es) % 16) return -1;
    if ((g.a_batch_stride * es) % 16 || (g.b_batch_stride * es) % 16) return -1;
    const int bn = g.N >= 192 ? 256 : (g.N >= 96 ? 128 : 64);
    const int bk = (int)(BK_BYTES / es);
    const int nkb = (int)((g.K + bk - 1) / bk);
    const int64_t mt = (g.M + BM - 1) / BM, nt = (g.N + bn - 1) / bn;
    const int64_t total = mt * nt * g.batch;
    const int sms = dev.sm_count > 0 ? dev.sm_count : 148;
    if (total < sms || total >= 0x7fffffff || nkb < 4) return -1;      // needs at least one full wave to overlap anything
    CUtensorMap ta, tb;
    const int64_t a_batches = (g.batch + g.a_bcast - 1) / g.a_bcast;
    if (!encode_operand(&ta, g.A, g.type, g.K, g.M, g.lda, a_batches, g.a_batch_stride, BM)) return -1;
    if (!encode_operand(&tb, g.B, g.type, g.K, g.N, g.ldb, g.batch, g.b_batch_stride, (uint32_t)bn)) return -1;
    PGemmParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.D = g.D; kp.ldd = g.ldd; kp.d_batch_stride = g.d_batch_stride;
    kp.M = g.M; kp.N = g.N;
    kp.num_k_blocks = nkb;
    kp.ne12 = (int)g.batch; kp.r2 = (int)g.a_bcast; kp.r3 = 1;
    kp.bias = g.bias; kp.bias_mode = g.bias ? g.bias_mode : 0;
    kp.residual = g.residual; kp.ldr = g.ldr; kp.r_batch_stride = g.d_batch_stride;
    const int fmt = g.type == GGML_TYPE_F16 ? 0 : (g.type == GGML_TYPE_BF16 ? 1 : 2);
    cudaError_t e = cudaErrorInvalidValue;
#define LAUNCH_P(BN_)                                                                                       \
    do {                                                                                                    \
        if (fmt == 0) e = launch_persistent<BN_, 0>(s, sms, ta, tb, kp, (int)mt, (int)nt, (int)total);      \
        else if (fmt == 1) e = launch_persistent<BN_, 1>(s, sms, ta, tb, kp, (int)mt, (int)nt, (int)total); \
        else e = launch_persistent<BN_, 2>(s, sms, ta, tb, kp, (int)mt, (int)nt, (int)total);               \
    } while (0)
    if (bn == 256) LAUNCH_P(256);
    else if (bn == 128) LAUNCH_P(128);
    else LAUNCH_P(64);
#undef LAUNCH_P
    if (e != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return 1;
}
