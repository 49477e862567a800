// gemm_tc.cu -- the dense-contraction core of the backend: D[n][m] = sum_k A[m][k] * B[n][k]
// on Blackwell 5th-generation tensor cores.
//
//   * operands are K-major (exactly ggml's MUL_MAT contract: src0 = [K, M], src1 = [K, N], dst = [M, N] with M
//     fastest, ggml/src/ggml.c:3282 / ggml-cpu.c:1406), described to the hardware by two 4-D TMA tensor maps
//     (K, rows, batch2, batch3) so ggml's batch broadcast and strided (permuted) batches need no copies
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B, zero OOB fill) stages 128 x BK / BN x BK tiles into a
//     STAGES-deep shared-memory ring guarded by full/empty mbarriers
//   * one elected thread issues tcgen05.mma (kind::f16 for F16/BF16, kind::tf32 for F32) with the f32 accumulator
//     tile 128 x BN living in TMEM; tcgen05.commit releases ring slots and finally signals the epilogue
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b), apply bias / activation / residual and store coalesced
//     along M (TMEM lane == m == ggml's fastest dst axis)
//   * small-M/N, large-K problems (SD1.5: M=256, N=1280, K=11520 -> 20 tiles on 148 SMs) are split along K;
//     partial tiles go to a workspace and the LAST-arriving CTA of each tile reduces them in split order
//     (deterministic) and runs the epilogue
//
// Roofline: tensor pipe.  Algorithmic work = 2*M*N*K flop per launch; the 128 x BN x 64 k-block costs 2*BN cycles
// of MMA issue at cta_group::1 (B300_MICROARCH "tcgen05 floor"), i.e. 8192 flop/cycle/SM.
#include "../b200_ops.h"
#include "b200_launch.cuh"
#include "sm100_ptx.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <string>
#include <cstring>
#include <algorithm>

using namespace sm100;

namespace {

constexpr int BM = 128;            // UMMA M (TMEM lanes)
constexpr int BK_BYTES = 128;      // one 128-byte swizzle atom of K per stage row
constexpr int A_STAGE_BYTES = BM * BK_BYTES;

struct GemmKParams {
    float* D;
    int64_t ldd, d_batch_stride;
    int64_t M, N;
    int num_k_blocks;
    int splits;
    int ne12;          // batch = i2 + ne12 * i3
    int r2, r3;        // A batch index = (i2 / r2, i3 / r3)
    const float* bias;
    int bias_mode;     // 0 none, 1 per m, 2 per n
    const float* residual;
    int64_t ldr, r_batch_stride;
    int act;           // 0 none, 1 silu, 2 gelu(tanh)
    const float* gate; // per-m factor after the activation, before the residual (b200_gemm_args::gate; always with a residual)
    // implicit-GEMM convolution (conv != 0): A is an NHWC f16 image read through a 4-D map (C, W, H, N) with halo boxes;
    // k-block kb -> filter tap kb / cblocks and 64-channel block kb % cblocks; rows of the tile are output pixels
    int conv;
    int conv_W;        // output width (== input width: stride 1, "same" padding)
    int conv_KW;       // filter width (taps = KH * KW)
    int conv_cblocks;  // IC / 64
    int conv_pad, conv_dil;
    int early;         // operands that may be fetched before griddepcontrol.wait (bit 0 = A, bit 1 = B): constant weights
    const char* wpf;   // weight operand to request from L2 up front (b200_gemm_args::wprefetch), null: off
    int64_t wpf_ld, wpf_rows, wpf_kbytes;
    int wpf_is_a;
    // optional 16-bit copy of the result, D's element layout (b200_gemm_args::D16): no residual, no activation on this kernel
    void* D16;
    int d16_bf16, skip_f32;
    unsigned long long* trace;   // optional: CTA (0,0,0) writes %globaltimer at its phase boundaries (tools/gemm_bench only)
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(slot)                                                                                    \
    do {                                                                                               \
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.trace[slot] = gtimer(); \
    } while (0)

template <int BN> struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK_BYTES;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    // The main loop is a latency ring (TMA issue -> data -> MMA -> commit -> slot free is ~1 us round trip), so depth is what
    // buys bandwidth: use (almost) all 227 KB -- 8 x 24 KB, 6 x 32 KB, 4 x 48 KB -- one CTA per SM.
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // + alignment slack
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ float epilogue_act(float v, int act) {
    if (act == 1) return v / (1.0f + expf(-v));
    if (act == 2) return 0.5f * v * (1.0f + tanhf(0.79788456080286535587989211986876f * v * (1.0f + 0.044715f * v * v)));
    return v;
}

__device__ __forceinline__ uint16_t round16(float v, int bf16) {
    if (bf16) { const __nv_bfloat16 h = __float2bfloat16_rn(v); return *(const uint16_t*)&h; }
    const __half h = __float2half_rn(v);
    return *(const uint16_t*)&h;
}

// FMT: 0 = f16, 1 = bf16, 2 = tf32 (operand format field of the instruction descriptor)
template <int BN, int FMT>
__global__ void __launch_bounds__(192, 1) k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    const GemmKParams p) {
    using C = Cfg<BN>;
    constexpr int BK = FMT == 2 ? 32 : 64;     // elements per 128-byte row
    constexpr int UMMA_K = FMT == 2 ? 8 : 16;  // 32 bytes of K per instruction
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[C::STAGES], empty_bar[C::STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int split = p.splits > 1 ? (int)cluster_ctarank() : 0;   // cluster (1,1,splits): rank == blockIdx.z % splits
    const int batch = blockIdx.z / p.splits;
    const int i2 = batch % p.ne12, i3 = batch / p.ne12;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kb0 = (int)(((int64_t)split * p.num_k_blocks) / p.splits);
    const int kb1 = (int)(((int64_t)(split + 1) * p.num_k_blocks) / p.splits);
    const int nkb = kb1 - kb0;
    if (threadIdx.x == 0) TRACE(0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (threadIdx.x == 0) TRACE(1);
    // Everything above is on-chip setup (barriers, TMEM, descriptor prefetch): under PDL it overlaps the predecessor's tail.
    // So does the first ring-full of the CONSTANT operand (p.early: bit 0 = A, bit 1 = B are model weights that no kernel of this
    // graph writes): the producer thread issues those TMA loads BEFORE griddepcontrol.wait, which hides the cold-HBM latency of the
    // weight stream behind the previous kernel; the activation operand follows after the wait.
    auto issue = [&](int i, int which) {
        const int s = i % C::STAGES;
        uint8_t* sa = smem + s * C::STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        const int k = (kb0 + i) * BK;
        if (p.conv) {
            const int kb = kb0 + i;
            const int tap = kb / p.conv_cblocks, cb = kb - tap * p.conv_cblocks;
            const int kh = tap / p.conv_KW, kw = tap - kh * p.conv_KW;
            const int y0 = m0 / p.conv_W, x0 = m0 - y0 * p.conv_W;
            // the box {64 ch, BW, BH} lands as 128 rows of 128 bytes: row = by * BW + bx == pixel m0 + row; out-of-image
            // (halo / tail) coordinates are zero-filled by the TMA unit, which is exactly the conv's zero padding
            if (which & 1) tma_load_4d(sa, &tmA, &full_bar[s], cb * 64, x0 + kw * p.conv_dil - p.conv_pad, y0 + kh * p.conv_dil - p.conv_pad, i2);
            if (which & 2) tma_load_4d(sb, &tmB, &full_bar[s], k, n0, 0, 0);
        } else {
            if (which & 1) tma_load_4d(sa, &tmA, &full_bar[s], k, m0, i2 / p.r2, i3 / p.r3);
            if (which & 2) tma_load_4d(sb, &tmB, &full_bar[s], k, n0, i2, i3);
        }
    };
    if (p.wpf) {
        // constant weights: request this CTA's slab (its rows, its K range) from L2 before the predecessor kernel has finished
        const int row0 = p.wpf_is_a ? m0 : n0;
        const int nrows = p.wpf_is_a ? BM : BN;
        const int64_t koff = (int64_t)kb0 * BK_BYTES;
        const int64_t kbytes = min((int64_t)nkb * BK_BYTES, p.wpf_kbytes - koff) & ~(int64_t)15;
        if (kbytes > 0)
            for (int r = threadIdx.x; r < nrows; r += 192)
                if (row0 + r < p.wpf_rows) l2_prefetch_bulk(p.wpf + (int64_t)(row0 + r) * p.wpf_ld + koff, (unsigned)kbytes);
    }
    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        const int pre = p.early ? min(nkb, C::STAGES) : 0;
        for (int i = 0; i < pre; ++i) {
            mbar_expect_tx(&full_bar[i], C::STAGE_BYTES);     // first pass over the ring: every slot is free
            issue(i, p.early);
        }
        pdl_wait();
        pdl_launch_dependents();
        for (int i = 0; i < pre; ++i) issue(i, 3 & ~p.early);
        for (int i = pre; i < nkb; ++i) {
            const int s = i % C::STAGES;
            const uint32_t ph = (i / C::STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_expect_tx(&full_bar[s], C::STAGE_BYTES);
            issue(i, 3);
        }
    } else {
        pdl_wait();
        pdl_launch_dependents();
    }

    if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc(FMT, BM, BN);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % C::STAGES;
            const uint32_t ph = (i / C::STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (lane == 0) {
                if (i == 0) TRACE(2);
                if (i == nkb - 1) TRACE(3);
                const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
                const uint64_t da = make_smem_desc_sw128(sa);
                const uint64_t db = make_smem_desc_sw128(sa + A_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // advance 32 bytes of K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
                    if (FMT == 2) mma_tf32(tmem_base, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    else mma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                mma_commit(&empty_bar[s]);                       // ring slot reusable once these MMAs retire
                if (i == nkb - 1) mma_commit(&tmem_full_bar);    // accumulator complete
            }
            __syncwarp();
        }
    } else if (warp >= 2) {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                       // TMEM lane quadrant this warp may access
        const int ml = q * 32 + lane;                 // row inside the tile
        const int64_t m = (int64_t)m0 + ml;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 64) TRACE(4);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool direct = p.splits == 1;
        float* Dp = p.D + (int64_t)batch * p.d_batch_stride;
        const float* Rp = p.residual ? p.residual + (int64_t)batch * p.r_batch_stride : nullptr;
        const float bias_m = (p.bias_mode == 1 && m < p.M) ? p.bias[m] : 0.f;
        const float gate_m = (p.gate && m < p.M) ? p.gate[m] : 1.f;
        float* sred = (float*)smem;   // [BN][BM] f32 partial tile; the operand ring is dead once tmem_full has fired
        const int ncols = (int)min((int64_t)BN, p.N - n0);
        const bool mvalid = m < p.M;
        if (!direct) {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) sred[(c0 + j) * BM + ml] = __uint_as_float(r[j]);
            }
        } else if (p.act == 0) {
            // fast path: every per-tile decision is hoisted; per element only add(s) + one predicated coalesced store
            float* dptr = Dp + (int64_t)n0 * p.ldd + m;
            const float* rptr = Rp ? Rp + (int64_t)n0 * p.ldr + m : nullptr;
            const float* bias_n = p.bias_mode == 2 ? p.bias + n0 : nullptr;
#pragma unroll 1
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                const float bn = (bias_n && c0 + lane < ncols) ? bias_n[c0 + lane] : 0.f;   // lane j carries the bias of column c0 + j
                tmem_ld_wait();
                if (p.D16) {
                    uint16_t* hptr = (uint16_t*)p.D16 + (int64_t)batch * p.d_batch_stride + (int64_t)n0 * p.ldd + m;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(r[j]) + bias_m + __shfl_sync(0xffffffffu, bn, j);
                        if (mvalid && c0 + j < ncols) {
                            if (!p.skip_f32) dptr[(int64_t)(c0 + j) * p.ldd] = v;
                            hptr[(int64_t)(c0 + j) * p.ldd] = round16(v, p.d16_bf16);
                        }
                    }
                } else if (rptr == nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(r[j]) + bias_m + __shfl_sync(0xffffffffu, bn, j);
                        if (mvalid && c0 + j < ncols) dptr[(int64_t)(c0 + j) * p.ldd] = v;
                    }
                } else {
                    // residual: issue all 32 loads before the first store (the compiler must assume rptr may alias dptr -- it often
                    // does, for in-place adds -- and would otherwise serialise load -> store -> load)
                    float rr[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) rr[j] = (mvalid && c0 + j < ncols) ? rptr[(int64_t)(c0 + j) * p.ldr] : 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float v = __uint_as_float(r[j]) + bias_m + __shfl_sync(0xffffffffu, bn, j);
                        if (p.gate) v = __fmul_rn(v, gate_m);
                        v += rr[j];
                        if (mvalid && c0 + j < ncols) dptr[(int64_t)(c0 + j) * p.ldd] = v;
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int64_t n = (int64_t)n0 + c0 + j;
                    if (m < p.M && n < p.N) {
                        float v = __uint_as_float(r[j]) + bias_m;
                        if (p.bias_mode == 2) v += p.bias[n];
                        v = epilogue_act(v, p.act);
                        if (p.gate) v = __fmul_rn(v, gate_m);
                        if (Rp) v += Rp[n * p.ldr + m];
                        Dp[n * p.ldd + m] = v;
                    }
                }
            }
        }
        tc_fence_before();
        if (threadIdx.x == 64) TRACE(5);
    }
    __syncthreads();
    if (p.splits > 1) {
        // ---- split-K reduction inside the thread-block cluster: the `splits` CTAs of one output tile form a cluster
        //      (1,1,splits); each keeps its partial tile in its own shared memory and reduces every splits-th group of 4 columns
        //      by reading the peers' tiles through distributed shared memory, in rank order (deterministic), then runs the
        //      epilogue for those columns.  No global workspace, no atomics.
        cluster_sync_all();
        if (threadIdx.x == 0) TRACE(6);
        if (warp >= 2) {
            // rank r owns the contiguous column range [r * BN / S, (r + 1) * BN / S); warp w takes every 4th column of it;
            // a lane reads 4 consecutive rows (16 bytes) of that column from each peer: one warp request = one 512-byte column
            const int w4 = warp - 2;
            const int64_t mrow = (int64_t)m0 + 4 * lane;
            float* Dp = p.D + (int64_t)batch * p.d_batch_stride;
            const float* Rp = p.residual ? p.residual + (int64_t)batch * p.r_batch_stride : nullptr;
            float4 bm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias_mode == 1) {
                if (mrow + 0 < p.M) bm.x = p.bias[mrow + 0];
                if (mrow + 1 < p.M) bm.y = p.bias[mrow + 1];
                if (mrow + 2 < p.M) bm.z = p.bias[mrow + 2];
                if (mrow + 3 < p.M) bm.w = p.bias[mrow + 3];
            }
            float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
            if (p.gate) {
                if (mrow + 0 < p.M) gm.x = p.gate[mrow + 0];
                if (mrow + 1 < p.M) gm.y = p.gate[mrow + 1];
                if (mrow + 2 < p.M) gm.z = p.gate[mrow + 2];
                if (mrow + 3 < p.M) gm.w = p.gate[mrow + 3];
            }
            const uint32_t sred_local = smem_u32(smem);
            uint32_t peer[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) peer[s] = s < p.splits ? dsmem_map(sred_local, (uint32_t)s) : 0u;
            const int ncols = (int)min((int64_t)BN, p.N - n0);
            const int cbeg = (split * BN) / p.splits, cend = min(((split + 1) * BN) / p.splits, ncols);
            const bool vec_ok = (mrow + 3 < p.M) && ((p.ldd & 3) == 0) && ((((uintptr_t)Dp) & 15) == 0) && ((m0 & 3) == 0);
#pragma unroll 1
            for (int c = cbeg + w4; c < cend; c += 4) {
                const uint32_t off = (uint32_t)(c * BM + 4 * lane) * 4u;
                float4 part[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) part[s] = s < p.splits ? dsmem_ld_f32x4(peer[s] + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int s = 0; s < 8; ++s) { v.x += part[s].x; v.y += part[s].y; v.z += part[s].z; v.w += part[s].w; }
                const int64_t n = (int64_t)n0 + c;
                const float bn = p.bias_mode == 2 ? p.bias[n] : 0.f;
                v.x += bm.x + bn; v.y += bm.y + bn; v.z += bm.z + bn; v.w += bm.w + bn;
                if (p.act) { v.x = epilogue_act(v.x, p.act); v.y = epilogue_act(v.y, p.act); v.z = epilogue_act(v.z, p.act); v.w = epilogue_act(v.w, p.act); }
                if (p.gate) { v.x = __fmul_rn(v.x, gm.x); v.y = __fmul_rn(v.y, gm.y); v.z = __fmul_rn(v.z, gm.z); v.w = __fmul_rn(v.w, gm.w); }
                float* dst = Dp + n * p.ldd + mrow;
                if (p.D16) {        // (launcher: no residual with a 16-bit copy)
                    uint16_t* h16 = (uint16_t*)p.D16 + (int64_t)batch * p.d_batch_stride + n * p.ldd + mrow;
                    if (vec_ok && (((uintptr_t)p.D16 | (uintptr_t)(p.d_batch_stride * 2)) & 7) == 0) {
                        if (!p.skip_f32) *(float4*)dst = v;
                        uint2 h;
                        h.x = (uint32_t)round16(v.x, p.d16_bf16) | ((uint32_t)round16(v.y, p.d16_bf16) << 16);
                        h.y = (uint32_t)round16(v.z, p.d16_bf16) | ((uint32_t)round16(v.w, p.d16_bf16) << 16);
                        *(uint2*)h16 = h;
                    } else {
                        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (mrow + u < p.M) {
                                if (!p.skip_f32) dst[u] = vv[u];
                                h16[u] = round16(vv[u], p.d16_bf16);
                            }
                    }
                } else if (vec_ok && (Rp == nullptr || (((p.ldr & 3) == 0) && ((((uintptr_t)Rp) & 15) == 0)))) {
                    if (Rp) { const float4 rr = *(const float4*)(Rp + n * p.ldr + mrow); v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
                    *(float4*)dst = v;
                } else {
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (mrow + u < p.M) dst[u] = vv[u] + (Rp ? Rp[n * p.ldr + mrow + u] : 0.f);
                }
            }
        }
        cluster_sync_all();   // nobody may exit (and release its shared memory) while a peer is still reading it
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
        if (lane == 0) TRACE(7);
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map cache + launch heuristics
// ------------------------------------------------------------------------------------------------
struct MapKey {
    const void* ptr;
    int type;
    uint64_t dims[4];
    uint64_t strides[3];
    uint32_t box_rows;
    bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        const uint64_t* w = (const uint64_t*)&k;
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
static_assert(sizeof(MapKey) % 8 == 0, "MapKey must be padded to 8 bytes");

std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

bool make_operand_map(CUtensorMap* out, const void* ptr, int type, int64_t K, int64_t rows, int64_t ld_elems, int64_t b2, int64_t b2_stride,
                      int64_t b3, int64_t b3_stride, uint32_t box_rows) {
    const int64_t es = type == GGML_TYPE_F32 ? 4 : 2;
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = ptr;
    key.type = type;
    key.dims[0] = (uint64_t)K; key.dims[1] = (uint64_t)rows; key.dims[2] = (uint64_t)b2; key.dims[3] = (uint64_t)b3;
    key.strides[0] = (uint64_t)(ld_elems * es);
    key.strides[1] = (uint64_t)(b2_stride * es);
    key.strides[2] = (uint64_t)(b3_stride * es);
    key.box_rows = box_rows;
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        auto it = g_map_cache.find(key);
        if (it != g_map_cache.end()) { *out = it->second; return true; }
    }
    auto enc = b200_get_tensormap_encoder();
    if (!enc) return false;
    CUtensorMapDataType dt = type == GGML_TYPE_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                   : (type == GGML_TYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    cuuint64_t dims[4] = {key.dims[0], key.dims[1], key.dims[2], key.dims[3]};
    cuuint64_t strides[3] = {key.strides[0], key.strides[1], key.strides[2]};
    // degenerate batch dims still need legal (16-byte multiple, non-zero) strides
    if (b2 == 1 || strides[1] == 0) strides[1] = strides[0] * dims[1];
    if (b3 == 1 || strides[2] == 0) strides[2] = strides[1] * dims[2];
    cuuint32_t box[4] = {(cuuint32_t)(BK_BYTES / es), box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    std::lock_guard<std::mutex> lock(g_map_mutex);
    if (g_map_cache.size() > 65536) g_map_cache.clear();
    g_map_cache[key] = *out;
    return true;
}

struct Plan { int bn; int splits; };

Plan choose_plan(const b200_device_info& dev, const b200_gemm_args& g, int num_k_blocks, double* cycles = nullptr) {
    const int sms = dev.sm_count > 0 ? dev.sm_count : 148;
    const int64_t mt = (g.M + BM - 1) / BM;
    double best = 1e30;
    Plan bestp{128, 1};
    const int bns[3] = {256, 128, 64};
    for (int bi = 0; bi < 3; ++bi) {
        const int bn = bns[bi];
        if (bn > 64 && g.N <= bn / 2) continue;                    // do not waste more than half of the N tile
        const int64_t nt = (g.N + bn - 1) / bn;
        const int64_t tiles = mt * nt * g.batch;
        for (int splits = 1; splits <= 8; ++splits) {   // portable cluster size limit
            if (splits > num_k_blocks) break;
            if (splits > 1 && (num_k_blocks / splits) < 4) break;   // keep the main loop meaningful
            const int64_t ctas = tiles * splits;
            const double per_sm = (double)((ctas + sms - 1) / sms);
            const double kb = (double)((num_k_blocks + splits - 1) / splits);
            // Measured model (tools/gemm_bench, B200): a 1-CTA main loop is bound by the SM's ingest port (~43-51 B/clk from L2 through TMA,
            // profiles/r02_gemm_model.md), not by the MMA: (128 + bn) x 128 B per k-block; the shared L2 caps the sum over active SMs.
            // (Constants as tuned in round 1 against this kernel's own sweep; the round-2 recalibration to a flat 43 B/clk picked worse
            // split-K plans for the 16x16 / 8x8 levels and was backed out.)
            const double active = (double)(ctas < sms ? ctas : sms);
            const double ingest = std::min(64.0 * 0.8, 4700.0 / active);
            const double kb_cycles = std::max(2.0 * bn, (128.0 + bn) * 128.0 / ingest);
            // fixed: setup + first data + accumulator hand-off + teardown ~ 3500 clk; epilogue ~ 35 clk per column (direct) or
            // smem staging + cluster barrier + DSMEM reduce of bn / splits columns
            double cta_cycles = kb * kb_cycles + 3500.0 + (splits > 1 ? 8.0 * bn + 2500.0 + 80.0 * bn / splits : 35.0 * bn);
            double t = per_sm * cta_cycles;
            if (t < best) { best = t; bestp = Plan{bn, splits}; }
        }
    }
    if (cycles) *cycles = best;
    return bestp;
}

// CTA-pair kernel (gemm_tc2.cu): GGML_B200_GEMM2 = 0 never, 1 when its modelled time beats the one-CTA plan (default), 2 whenever legal
int gemm2_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GGML_B200_GEMM2");
        v = (e && *e) ? atoi(e) : 1;      // validated on B200: 275 / 275 shape x plan combinations element-exact vs the one-CTA kernel, full GPU suite green when forced
    }
    return v;
}

int gemm_log() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GGML_B200_GEMM_LOG"); v = (e && *e) ? atoi(e) : 0; }
    return v;
}

struct Plan2 { int bn; int splits; double cycles; };

Plan2 choose_plan2(const b200_device_info& dev, int64_t M, int64_t N, int64_t batch, int nkb) {
    const int sms = dev.sm_count > 0 ? dev.sm_count : 148;
    Plan2 best{0, 1, 1e30};
    const int bns[] = {256, 224, 192, 160, 128, 96, 64, 48, 32};
    // tuning overrides (scripts/tune_gemm.py): restrict the candidates to one tile width / split factor
    static int force_bn = -1, force_splits = -1;
    if (force_bn < 0) { const char* e = getenv("GGML_B200_GEMM2_BN"); force_bn = (e && *e) ? atoi(e) : 0; }
    if (force_splits < 0) { const char* e = getenv("GGML_B200_GEMM2_SPLITS"); force_splits = (e && *e) ? atoi(e) : 0; }
    for (int bn : bns) {
        if (force_bn && bn != force_bn && !(N <= force_bn / 2 && bn < force_bn)) continue;
        if (bn > 32 && N <= bn / 2) continue;
        const int64_t tiles = ((M + 255) / 256) * ((N + bn - 1) / bn) * batch;
        for (int splits = 1; splits <= 4; splits *= 2) {
            if (splits > 1 && (tiles * 2 * splits > sms || nkb / splits < 4)) break;
            if (force_splits && splits != force_splits && !(splits == 1 && (tiles * 2 * force_splits > sms || nkb / force_splits < 4))) continue;
            const double t = b200_gemm_tc2_model(dev, M, N, batch, nkb, bn, splits);
            if (t < best.cycles) best = Plan2{bn, splits, t};
        }
    }
    return best;
}

// halo-reuse convolution (gemm_tc2.cu).  Its own cost model, fitted to the hardware sweep of tools/gemm_bench (GEMM_BENCH_CONV=1: ten
// UNet / VAE / SDXL 3x3 convolutions x tile width x split-K x taps per stage, profiles/r02_conv_halo_sweep.log; 12 % rms, picks within
// 0-14 % of the best measured plan for every shape).  What the sweep showed and the GEMM model above does not capture:
//   * a k-block costs max(2.64 bn, staged bytes / 31.8, (operand reads + staged bytes) / 97.8) clk: besides the MMA issue rate and the
//     L2 -> SM fill, the SHARED-MEMORY port is a limit of its own -- every K = 16 MMA of the pair re-reads 128 rows of A and bn / 2 rows of
//     B (32 B each) per CTA, the TMA fill writes through the same port (narrow tiles re-read A per few columns: bn = 64 is port bound)
//   * the epilogue of a tile (~2800 clk) is NOT hidden behind the next tile's main loop in practice (it shares that port), and the
//     fixed cost of a launch (prologue, first fill, last drain) is ~13.9 k clk
// Halo plans beat every per-tap plan of the same shape in the sweep, so when the halo envelope holds the per-tap plan is not considered.
struct Plan2H { int bn; int splits; int taps; double cycles; };
double conv_halo_model(const b200_device_info& dev, int64_t M, int64_t N, int64_t batch, int cblocks, int bn, int splits, int taps) {
    const int sms = dev.sm_count > 0 ? dev.sm_count : 148;
    const int64_t tiles = ((M + 255) / 256) * ((N + bn - 1) / bn) * batch;
    const int nkb = 9 * cblocks;
    const double kb = (double)((nkb + splits - 1) / splits);
    const double staged = (taps == 9 ? 23040.0 / 9.0 : 20480.0 / 3.0) + bn / 2.0 * 128.0;     // bytes per 64-wide k-block and CTA
    const double kb_cycles = std::max(std::max(2.64 * bn, staged / 31.8), (4.0 * (4096.0 + 16.0 * bn) + staged) / 97.8);
    if (splits == 1) {
        const int64_t pairs = std::min<int64_t>(tiles, sms / 2);
        return 13870.0 + (double)((tiles + pairs - 1) / pairs) * (kb * kb_cycles + 2800.0);
    }
    const int csize = 2 * splits;
    const int64_t max_clusters = csize == 4 ? 36 : (csize == 6 ? 20 : 13);   // concurrent clusters the GPCs (16-20 SMs each) can host
    return (double)((tiles + max_clusters - 1) / max_clusters) * (13870.0 + kb * kb_cycles);
}
Plan2H choose_plan2_halo(const b200_device_info& dev, int64_t M, int64_t N, int64_t batch, int cblocks) {
    const int sms = dev.sm_count > 0 ? dev.sm_count : 148;
    Plan2H best{0, 1, 0, 1e30};
    const int bns[] = {256, 192, 160, 128, 96, 64, 48, 32};      // (the widths the sweep covered; 48 / 32 only ever for N <= 32)
    for (int bn : bns) {
        if (bn > 32 && N <= bn / 2) continue;
        if (bn < 64 && N > 32) continue;
        const int64_t tiles = ((M + 255) / 256) * ((N + bn - 1) / bn) * batch;
        for (int splits = 1; splits <= 4; splits *= 2) {
            const int taps = b200_conv_tc2_halo_taps(bn, splits);
            if (!taps) continue;
            const int nst = taps == 9 ? cblocks : 3 * cblocks;          // ring stages per tile: what split-K divides
            if (splits > 1 && (tiles * 2 * splits > sms || nst / splits < 2)) break;
            const double t = conv_halo_model(dev, M, N, batch, cblocks, bn, splits, taps);
            if (t < best.cycles) best = Plan2H{bn, splits, taps, t};
        }
    }
    return best;
}

template <int BN, int FMT>
cudaError_t launch_cfg(cudaStream_t s, dim3 grid, const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& kp) {
    using C = Cfg<BN>;
    static bool configured[B200_MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    cfg.numAttrs = b200_launch_attrs(attr, (unsigned)kp.splits);
    cfg.attrs = attr;
    return cudaLaunchKernelEx(&cfg, k_gemm_tc<BN, FMT>, ta, tb, kp);
}

}  // namespace

// Diagnostic C-ABI (include/ggml-b200.h): the plan the halo-reuse convolution front end would choose for a 3x3 convolution of `batch`
// images of H x W x C -> OC on a device with `sm_count` SMs.  Pure host code: lets a CPU test pin the fitted cost model against the
// committed hardware sweep (profiles/r02_conv_halo_sweep.log).  Returns 0 when the shape is outside the halo envelope.
extern "C" int ggml_backend_b200_debug_conv_plan(int64_t batch, int64_t H, int64_t W, int64_t C, int64_t OC, int sm_count, int* bn, int* splits, int* taps,
                                                 double* model_us) {
    if (batch < 1 || H < 1 || W < 1 || C % 64 || C < 64 || OC < 1 || W % 8 || H % 16 || (H * W) % 128) return 0;
    b200_device_info dev;
    memset(&dev, 0, sizeof(dev));
    dev.sm_count = sm_count;
    const Plan2H ph = choose_plan2_halo(dev, H * W, OC, batch, (int)(C / 64));
    if (ph.bn <= 0) return 0;
    if (bn) *bn = ph.bn;
    if (splits) *splits = ph.splits;
    if (taps) *taps = ph.taps;
    if (model_us) *model_us = ph.cycles / 1965.0;
    return 1;
}

size_t b200_gemm_tc_workspace_bytes(const b200_device_info&, const b200_gemm_args&) {
    return 0;   // split-K partials live in distributed shared memory
}

int b200_launch_gemm_tc(cudaStream_t s, const b200_device_info& dev, const b200_gemm_args& g, void* workspace, size_t workspace_bytes) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return 0;
    const int64_t es = g.type == GGML_TYPE_F32 ? 4 : 2;
    if (g.type != GGML_TYPE_F32 && g.type != GGML_TYPE_F16 && g.type != GGML_TYPE_BF16) return -1;
    // TMA legality: 16-byte aligned bases and strides
    if (((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15)) return -1;
    if ((g.lda * es) % 16 || (g.ldb * es) % 16) return -1;
    if (g.K <= 0) return -1;
    if (g.gate && !g.residual) return -1;       // the gated epilogue exists on the residual paths only
    const int bk = (int)(BK_BYTES / es);
    const int nkb = (int)((g.K + bk - 1) / bk);
    if (g.geglu) {
        // GEGLU projection: exists on the CTA-pair kernel only (one tile per pair step, no split-K); widest-to-narrowest tile by the model
        if (!gemm2_mode() || g.type == GGML_TYPE_F32) return -1;
        int best_bn = 0;
        double best = 1e30;
        const int bns[] = {256, 192, 160, 128, 96, 64};
        for (int bn : bns) {
            if (bn > 64 && g.N <= bn / 2) continue;
            const double t = b200_gemm_tc2_model(dev, g.M, g.N, g.batch, nkb, bn, 1);
            if (t < best) { best = t; best_bn = bn; }
        }
        if (!best_bn || b200_launch_gemm_tc2(s, dev, g, best_bn, 1) <= 0) return -1;
        if (gemm_log()) fprintf(stderr, "GEMMLOG pair geglu M %lld N %lld K %lld batch %lld bn %d\n", (long long)g.M, (long long)g.N, (long long)g.K, (long long)g.batch, best_bn);
        return 2;
    }
    double cycles1 = 0;
    Plan pl = choose_plan(dev, g, nkb, &cycles1);
    if (gemm2_mode() && g.type != GGML_TYPE_F32 && !g.trace && !g.early && g.M > BM) {
        const Plan2 p2 = choose_plan2(dev, g.M, g.N, g.batch, nkb);
        if (p2.bn > 0 && (gemm2_mode() >= 2 || p2.cycles < cycles1)) {
            const int r = b200_launch_gemm_tc2(s, dev, g, p2.bn, p2.splits);
            if (r > 0) {
                if (gemm_log()) fprintf(stderr, "GEMMLOG pair gemm M %lld N %lld K %lld batch %lld bn %d splits %d model1 %.0f model2 %.0f\n", (long long)g.M, (long long)g.N, (long long)g.K, (long long)g.batch, p2.bn, p2.splits, cycles1, p2.cycles);
                return 2;       // 2: launched on the CTA-pair kernel
            }
        }
    }

    // batch decomposition: args carry a flat batch with an A broadcast ratio; map to (i2, i3) = (batch, 1)
    CUtensorMap ta, tb;
    const int64_t a_batches = (g.batch + g.a_bcast - 1) / g.a_bcast;
    if ((g.a_batch_stride * es) % 16 || (g.b_batch_stride * es) % 16) return -1;
    if (!make_operand_map(&ta, g.A, g.type, g.K, g.M, g.lda, a_batches, g.a_batch_stride, 1, 0, BM)) return -1;
    if (!make_operand_map(&tb, g.B, g.type, g.K, g.N, g.ldb, g.batch, g.b_batch_stride, 1, 0, (uint32_t)pl.bn)) return -1;

    GemmKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.D = g.D; kp.ldd = g.ldd; kp.d_batch_stride = g.d_batch_stride;
    kp.M = g.M; kp.N = g.N;
    kp.num_k_blocks = nkb;
    kp.splits = pl.splits;
    kp.ne12 = (int)g.batch; kp.r2 = (int)g.a_bcast; kp.r3 = 1;
    kp.bias = g.bias; kp.bias_mode = g.bias ? g.bias_mode : 0;
    kp.residual = g.residual; kp.ldr = g.ldr; kp.r_batch_stride = g.d_batch_stride;
    kp.act = g.act;
    kp.gate = g.gate;
    if ((g.wprefetch & 1) && a_batches == 1) { kp.wpf = (const char*)g.A; kp.wpf_ld = g.lda * es; kp.wpf_rows = g.M; kp.wpf_kbytes = g.K * es; kp.wpf_is_a = 1; }
    else if ((g.wprefetch & 2) && g.batch == 1) { kp.wpf = (const char*)g.B; kp.wpf_ld = g.ldb * es; kp.wpf_rows = g.N; kp.wpf_kbytes = g.K * es; kp.wpf_is_a = 0; }
    kp.early = g.early & 3;
    kp.trace = (unsigned long long*)g.trace;
    if (g.D16 && !g.residual && g.act == 0 && (g.d16_type == GGML_TYPE_F16 || g.d16_type == GGML_TYPE_BF16) && !((uintptr_t)g.D16 & 1)) {
        kp.D16 = g.D16; kp.d16_bf16 = g.d16_type == GGML_TYPE_BF16; kp.skip_f32 = g.skip_f32;
    } else if (g.D16 && g.d16_strict) {
        return -1;
    }
    const int64_t mt = (g.M + BM - 1) / BM, nt = (g.N + pl.bn - 1) / pl.bn;
    if (mt > 0x7fffffff || nt > 65535 || g.batch * pl.splits > 65535) return -1;
    (void)workspace; (void)workspace_bytes;
    dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)(g.batch * pl.splits));
    if (gemm_log()) fprintf(stderr, "GEMMLOG one gemm M %lld N %lld K %lld batch %lld bn %d splits %d model1 %.0f\n", (long long)g.M, (long long)g.N, (long long)g.K, (long long)g.batch, pl.bn, pl.splits, cycles1);
    const int fmt = g.type == GGML_TYPE_F16 ? 0 : (g.type == GGML_TYPE_BF16 ? 1 : 2);
    cudaError_t e = cudaErrorInvalidValue;
#define LAUNCH(BN_)                                                          \
    do {                                                                     \
        if (fmt == 0) e = launch_cfg<BN_, 0>(s, grid, ta, tb, kp);           \
        else if (fmt == 1) e = launch_cfg<BN_, 1>(s, grid, ta, tb, kp);      \
        else e = launch_cfg<BN_, 2>(s, grid, ta, tb, kp);                    \
    } while (0)
    if (pl.bn == 256) LAUNCH(256);
    else if (pl.bn == 128) LAUNCH(128);
    else LAUNCH(64);
#undef LAUNCH
    if (e != cudaSuccess) {
        fprintf(stderr, "[ggml-b200] tcgen05 GEMM launch failed: %s\n", cudaGetErrorString(e));
        return -1;
    }
    if (kp.D16 && g.d16_done) *g.d16_done = 1;
    return 1;
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM convolution front end (stride 1, "same" padding, IC % 64 == 0, W | 128 or 128 | W)
// ------------------------------------------------------------------------------------------------
bool b200_conv_tc_supported(int64_t N, int64_t H, int64_t W, int64_t C, int64_t OC, int KH, int KW, int s0, int s1, int p0, int p1, int d0, int d1) {
    if (s0 != 1 || s1 != 1 || d0 != d1 || p0 != p1) return false;
    if (KH != KW || (KH != 1 && KH != 3)) return false;
    if (p0 != d0 * (KH - 1) / 2) return false;             // output size == input size
    if (C % 64 != 0 || C <= 0 || OC <= 0) return false;
    if (!((W <= 128 && 128 % W == 0) || (W % 128 == 0))) return false;
    if (N > 65535 || H * W > 0x7fffffff) return false;
    return true;
}

size_t b200_conv_tc_workspace_bytes(const b200_device_info& dev, const b200_conv_args& c) {
    b200_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.type = GGML_TYPE_F16; g.M = c.H * c.W; g.N = c.OC; g.K = (int64_t)c.KH * c.KW * c.C; g.batch = c.N; g.a_bcast = 1;
    return b200_gemm_tc_workspace_bytes(dev, g);
}

int b200_launch_conv_tc(cudaStream_t s, const b200_device_info& dev, const b200_conv_args& c, void* workspace, size_t workspace_bytes) {
    if (!b200_conv_tc_supported(c.N, c.H, c.W, c.C, c.OC, c.KH, c.KW, 1, 1, c.pad, c.pad, c.dil, c.dil)) return -1;
    if (((uintptr_t)c.x_nhwc & 15) || ((uintptr_t)c.w_packed & 15)) return -1;
    b200_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.type = GGML_TYPE_F16; g.M = c.H * c.W; g.N = c.OC; g.K = (int64_t)c.KH * c.KW * c.C; g.batch = c.N; g.a_bcast = 1;
    const int nkb = (int)(g.K / 64);
    double cycles1 = 0;
    Plan pl = choose_plan(dev, g, nkb, &cycles1);
    // (a peer destination c.D2 does not influence the choice: the serial and the split sampler must run the very same plans to stay
    //  bit-identical; when this launch ends up on the one-CTA kernel the executor pushes the tensor with kernels/peer.cu instead)
    if (gemm2_mode() && g.M > BM && g.M % 128 == 0) {
        Plan2 p2 = choose_plan2(dev, g.M, g.N, g.batch, nkb);
        int taps = 0;
        static int halo_en = -1;
        if (halo_en < 0) { const char* e = getenv("GGML_B200_CONV_HALO"); halo_en = (e && *e) ? atoi(e) : 1; }
        // A residual that does not fit L2 is read in the halo patches' 32-byte pieces (8 pixels of one row per channel) straight from HBM:
        // measured 221 us against 91 us for the same 512 x 512 x 128 convolution without one.  The per-tap kernel reads and writes 512
        // contiguous bytes per channel and row; it is the faster plan there (VAE 512 / 256 levels), the halo plan everywhere else.
        const bool big_residual = c.residual && (double)c.OC * (double)g.M * (double)c.N * 4.0 >= 48.0 * 1024 * 1024;
        if (halo_en && !big_residual && c.KH == 3 && c.KW == 3 && c.pad == 1 && c.dil == 1 && c.W % 8 == 0 && c.H % 16 == 0) {
            const Plan2H ph = choose_plan2_halo(dev, g.M, g.N, g.batch, (int)(c.C / 64));
            if (ph.bn > 0) { p2 = Plan2{ph.bn, ph.splits, std::min(ph.cycles, p2.bn > 0 ? p2.cycles : ph.cycles)}; taps = ph.taps; }
        }
        if (p2.bn > 0 && (gemm2_mode() >= 2 || p2.cycles < cycles1)) {
            int r = b200_launch_conv_tc2(s, dev, c, p2.bn, p2.splits, taps);
            if (r <= 0 && taps) {          // (outside the halo envelope after all: the per-tap plan)
                p2 = choose_plan2(dev, g.M, g.N, g.batch, nkb);
                taps = 0;
                if (p2.bn > 0 && (gemm2_mode() >= 2 || p2.cycles < cycles1)) r = b200_launch_conv_tc2(s, dev, c, p2.bn, p2.splits, 0);
            }
            if (r > 0) {
                if (gemm_log()) fprintf(stderr, "GEMMLOG pair conv M %lld N %lld K %lld batch %lld bn %d splits %d taps %d model1 %.0f model2 %.0f\n", (long long)g.M, (long long)g.N, (long long)g.K, (long long)g.batch, p2.bn, p2.splits, taps, cycles1, p2.cycles);
                return 2;
            }
        }
    }

    // A: NHWC image, 4-D (C, W, H, N); box = 64 channels x BW x BH pixels with BW * BH == 128
    const uint32_t BW = (uint32_t)(c.W < 128 ? c.W : 128), BH = 128 / BW;
    CUtensorMap ta, tb;
    {
        auto enc = b200_get_tensormap_encoder();
        if (!enc) return -1;
        cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
        cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
        cuuint32_t box[4] = {64, BW, BH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(c.x_nhwc), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
    }
    if (!make_operand_map(&tb, c.w_packed, GGML_TYPE_F16, g.K, c.OC, g.K, 1, 0, 1, 0, (uint32_t)pl.bn)) return -1;

    GemmKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.D = c.D; kp.ldd = c.H * c.W; kp.d_batch_stride = c.OC * c.H * c.W;
    kp.M = g.M; kp.N = g.N;
    kp.num_k_blocks = nkb;
    kp.splits = pl.splits;
    kp.ne12 = (int)c.N; kp.r2 = 1; kp.r3 = 1;
    if (c.w_prefetch) { kp.wpf = (const char*)c.w_packed; kp.wpf_ld = g.K * 2; kp.wpf_rows = c.OC; kp.wpf_kbytes = g.K * 2; kp.wpf_is_a = 0; }
    kp.bias = c.bias; kp.bias_mode = c.bias ? 2 : 0;
    kp.residual = c.residual; kp.ldr = c.H * c.W; kp.r_batch_stride = c.OC * c.H * c.W;
    kp.act = 0;
    kp.early = c.w_const ? 2 : 0;
    kp.conv = 1; kp.conv_W = (int)c.W; kp.conv_KW = c.KW; kp.conv_cblocks = (int)(c.C / 64); kp.conv_pad = c.pad; kp.conv_dil = c.dil;
    const int64_t mt = (g.M + BM - 1) / BM, nt = (g.N + pl.bn - 1) / pl.bn;
    if (nt > 65535 || g.batch * pl.splits > 65535) return -1;
    (void)workspace; (void)workspace_bytes;
    dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)(g.batch * pl.splits));
    if (gemm_log()) fprintf(stderr, "GEMMLOG one conv M %lld N %lld K %lld batch %lld bn %d splits %d model1 %.0f\n", (long long)g.M, (long long)g.N, (long long)g.K, (long long)g.batch, pl.bn, pl.splits, cycles1);
    cudaError_t e;
    if (pl.bn == 256) e = launch_cfg<256, 0>(s, grid, ta, tb, kp);
    else if (pl.bn == 128) e = launch_cfg<128, 0>(s, grid, ta, tb, kp);
    else e = launch_cfg<64, 0>(s, grid, ta, tb, kp);
    if (e != cudaSuccess) {
        fprintf(stderr, "[ggml-b200] implicit-GEMM conv launch failed: %s\n", cudaGetErrorString(e));
        return -1;
    }
    return 1;
}
