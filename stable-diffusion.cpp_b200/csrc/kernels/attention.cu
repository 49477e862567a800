// attention.cu -- fused FLASH_ATTN_EXT for Blackwell: softmax(Q K^T * scale + mask) V in one kernel, scores never
// leave the SM.  (ggml op: ggml/src/ggml.c:5476; oracle: ggml/src/ggml-cpu/ops.cpp:8468-9176.)
//
//   q f32 [d, Lq, H, N] (converted to f16 on load, like the oracle: ops.cpp:8586), k f16 [d, Lk, Hkv, N],
//   v^T f16 [Lk_pad, dv, Hkv, N] (packed by a pre-pass so that both MMA operands are K-major), mask f16 [Lk, >=Lq, ..] | null
//   dst f32 [dv, H, Lq, N]
//
// One CTA per (128-query tile, head, batch); 6 warps:
//   warp 0   TMA producer: K tiles [BLOCK_N x d] and V^T tiles [dv x BLOCK_N] into 2-deep rings (SWIZZLE_128B, zero OOB fill
//            supplies both the d -> multiple-of-16 padding and the Lk tail)
//   warp 1   tcgen05.mma issuer: S_j = Q K_j^T (128 x BLOCK_N, f32 in TMEM, double buffered) is issued one tile AHEAD of
//            O += P_j V_j, so the tensor pipe works on S_{j+1} while the softmax warps chew on S_j
//   warps 2-5  one query row per thread: tcgen05.ld S, running max / sum (exp2 with scale*log2e folded), P_j -> f16 into
//            shared memory in the UMMA K-major swizzled layout; O lives in TMEM across all KV tiles and is rescaled in place
//            (tcgen05.ld / st) only when a row maximum grows by more than 2^8 ("lazy rescaling"), final O / l -> global
//
// Roofline: tensor pipe, 4 * Lq * Lk * d flop per head (QK^T and PV).  HBM traffic is Q + K + V + O only.
#include "../b200_ops.h"
#include "b200_launch.cuh"
#include "sm100_ptx.cuh"

#include <cuda_fp16.h>
#include <cstring>
#include <mutex>
#include <unordered_map>

using namespace sm100;

namespace {

constexpr int BLOCK_M = 128;
constexpr float kLazyThreshold = 8.0f;   // log2 units

struct FaParams {
    const float* q;
    int64_t q_nb1, q_nb2, q_nb3;   // bytes
    float* dst;
    int64_t dst_nb1, dst_nb2, dst_nb3;   // bytes: head, query, batch
    __half* dst16;                       // optional f16 copy of dst, same element layout (byte strides dst_nb / 2): the operand of the
                                         // output projection that follows, written here instead of by a separate pack kernel
    const __half* mask;
    int64_t m_nb1, m_nb2, m_nb3;
    int m_ne2, m_ne3;
    int d, dv16, Lq, Lk, H, rk;    // rk = H / Hkv
    int v_mn;                      // V is read in its natural [key][dv] layout (MN-major B operand of the P.V product); 0: pre-transposed V^T
    int pos_scale;                 // scale > 0: the row maximum may be taken before scaling
    int q_vec;                     // Q rows are 16-byte aligned and d % 4 == 0: coalesced float4 loads
    int skip_f32;                  // only the f16 copy is wanted (its single reader is the output projection): no f32 stores
    float scale_log2;              // scale * log2(e)
    float log2e;
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NATOM, int BLOCK_N> struct FaCfg {
    static constexpr int Q_BYTES = NATOM * BLOCK_M * 128;
    static constexpr int K_STAGE = NATOM * BLOCK_N * 128;
    static constexpr int VROWS_MAX = NATOM * 64;                       // dv16 <= NATOM * 64
    static constexpr int V_STAGE = (BLOCK_N / 64) * VROWS_MAX * 128;
    static constexpr int P_BYTES = (BLOCK_N / 64) * BLOCK_M * 128;
    static constexpr int BAR_BYTES = 128;                               // 13 mbarriers + the TMEM base word, carved from the dynamic allocation
    // no alignment slack and no static shared memory: the dynamic window is 1024-byte aligned by declaration, and two CTAs of the
    // d = 128 / 64-key configuration fit one SM exactly (2 x (112 KB + 128 B) + the 1 KB the system reserves per CTA <= 228 KB)
    static constexpr int SMEM = Q_BYTES + 2 * K_STAGE + 2 * V_STAGE + P_BYTES + BAR_BYTES;
    static constexpr int MIN_CTAS = (2 * (SMEM + 1024) <= 228 * 1024 && 2 * ((2 * BLOCK_N + NATOM * 64) <= 256 ? 256 : 512) <= 512) ? 2 : 1;
    static constexpr int TMEM_COLS = (2 * BLOCK_N + NATOM * 64) <= 256 ? 256 : 512;
};

template <int NATOM, int BLOCK_N>
__global__ void __launch_bounds__(192, FaCfg<NATOM, BLOCK_N>::MIN_CTAS) k_flash_attn(const __grid_constant__ CUtensorMap tmK,
                                                                                     const __grid_constant__ CUtensorMap tmV, const FaParams p) {
    using C = FaCfg<NATOM, BLOCK_N>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + C::Q_BYTES;
    uint8_t* sV = sK + 2 * C::K_STAGE;
    uint8_t* sP = sV + 2 * C::V_STAGE;
    uint64_t* bars = (uint64_t*)(sP + C::P_BYTES);
    uint64_t* k_full = bars;            // [2]
    uint64_t* k_empty = bars + 2;       // [2]
    uint64_t* v_full = bars + 4;        // [2]
    uint64_t* v_empty = bars + 6;       // [2]
    uint64_t* s_full = bars + 8;        // [2]
    uint64_t& p_full = bars[10];
    uint64_t& pv_done = bars[11];
    uint64_t& q_ready = bars[12];
    uint32_t& tmem_base_smem = *(uint32_t*)(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BLOCK_M;
    const int h = blockIdx.y, nb = blockIdx.z;
    const int hkv = h / p.rk;
    const int nblk = (p.Lk + BLOCK_N - 1) / BLOCK_N;
    const int v_atom_bytes = p.dv16 * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1);
        }
        mbar_init(&p_full, 128);
        mbar_init(&pv_done, 1);
        mbar_init(&q_ready, 128);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    if (warp == 1) { tmem_alloc(&tmem_base_smem, C::TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t tmem_S0 = tmem_base, tmem_O = tmem_base + 2 * BLOCK_N;
    pdl_wait();                 // on-chip setup above overlaps the predecessor's tail (b200_launch.cuh)
    pdl_launch_dependents();

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            for (int j = 0; j < nblk; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&k_empty[s], ph ^ 1);
                mbar_expect_tx(&k_full[s], C::K_STAGE);
                for (int a = 0; a < NATOM; ++a)
                    tma_load_4d(sK + s * C::K_STAGE + a * (BLOCK_N * 128), &tmK, &k_full[s], a * 64, j * BLOCK_N, hkv, nb);
                mbar_wait(&v_empty[s], ph ^ 1);
                if (p.v_mn) {
                    // V as stored, [key][dv]: one box {64 dv, BLOCK_N keys} per 64-column block of dv (zero fill beyond dv and beyond Lk)
                    const int natom_v = (p.dv16 + 63) / 64;
                    mbar_expect_tx(&v_full[s], natom_v * (BLOCK_N * 128));
                    for (int a = 0; a < natom_v; ++a)
                        tma_load_4d(sV + s * C::V_STAGE + a * (BLOCK_N * 128), &tmV, &v_full[s], a * 64, j * BLOCK_N, hkv, nb);
                } else {
                    mbar_expect_tx(&v_full[s], (BLOCK_N / 64) * v_atom_bytes);
                    for (int a = 0; a < BLOCK_N / 64; ++a)
                        tma_load_4d(sV + s * C::V_STAGE + a * v_atom_bytes, &tmV, &v_full[s], j * BLOCK_N + a * 64, 0, hkv, nb);
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        const uint32_t idesc_qk = make_idesc(0, BLOCK_M, BLOCK_N);
        // P.V: B = V.  Pre-transposed V^T is K-major like every other operand; V in its natural layout is the MN-major form (bit 16 of the
        // instruction descriptor): rows of 64 dv values (128 B, swizzled), 8-key groups 1024 B apart (SBO), 64-column blocks of dv
        // BLOCK_N * 128 B apart (LBO) -- no transposition pass, no V^T workspace
        const uint32_t idesc_pv = make_idesc(0, BLOCK_M, (uint32_t)p.dv16) | (p.v_mn ? (1u << 16) : 0u);
        const int ksteps_qk = (p.d + 15) / 16;
        auto issue_qk = [&](int j) {
            const int s = j & 1;
            mbar_wait(&k_full[s], (j >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t qb = smem_u32(sQ), kb = smem_u32(sK + s * C::K_STAGE);
                for (int kk = 0; kk < ksteps_qk; ++kk) {
                    const uint64_t da = make_smem_desc_sw128(qb + (kk >> 2) * (BLOCK_M * 128) + (kk & 3) * 32);
                    const uint64_t db = make_smem_desc_sw128(kb + (kk >> 2) * (BLOCK_N * 128) + (kk & 3) * 32);
                    mma_f16(tmem_S0 + s * BLOCK_N, da, db, idesc_qk, kk > 0 ? 1u : 0u);
                }
                mma_commit(&s_full[s]);
                mma_commit(&k_empty[s]);
            }
            __syncwarp();
        };
        mbar_wait(&q_ready, 0);
        tc_fence_after();
        issue_qk(0);
        for (int j = 0; j < nblk; ++j) {
            if (j + 1 < nblk) issue_qk(j + 1);
            const int s = j & 1;
            mbar_wait(&p_full, j & 1);
            mbar_wait(&v_full[s], (j >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t pb = smem_u32(sP), vb = smem_u32(sV + s * C::V_STAGE);
#pragma unroll
                for (int kk = 0; kk < BLOCK_N / 16; ++kk) {
                    const uint64_t da = make_smem_desc_sw128(pb + (kk >> 2) * (BLOCK_M * 128) + (kk & 3) * 32);
                    uint64_t db;
                    if (p.v_mn) {
                        const uint32_t addr = vb + kk * 2048;                          // 16 keys = two 8-key groups of 1024 B
                        db = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((BLOCK_N * 128) >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
                             ((uint64_t)2 << 61);
                    } else {
                        db = make_smem_desc_sw128(vb + (kk >> 2) * v_atom_bytes + (kk & 3) * 32);
                    }
                    mma_f16(tmem_O, da, db, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                }
                mma_commit(&pv_done);
                mma_commit(&v_empty[s]);
            }
            __syncwarp();
        }
    } else {
        // ============================== softmax / correction / epilogue ==============================
        const int qd = warp & 3;
        const int r = qd * 32 + lane;              // row in the tile == TMEM lane
        const int qi = q0 + r;                     // global query index
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;

        // ---- Q tile: f32 global -> f16, K-major 128B-swizzled shared memory (zero padded to NATOM*64 columns)
        if (p.q_vec) {
            // coalesced: the warp walks its 32 rows; lane L converts columns 4L .. 4L+3 (+128 per pass) of the row -- one 512-byte row
            // segment per load instruction instead of 32 scattered 4-byte reads
#pragma unroll 1
            for (int rr = 0; rr < 32; ++rr) {
                const int row = qd * 32 + rr;
                const int qrow_i = q0 + row;
                const float* qrow = (const float*)((const char*)p.q + (int64_t)qrow_i * p.q_nb1 + (int64_t)h * p.q_nb2 + (int64_t)nb * p.q_nb3);
#pragma unroll
                for (int c0 = 0; c0 < NATOM * 64; c0 += 128) {
                    const int col = c0 + lane * 4;
                    if (col < NATOM * 64) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (qrow_i < p.Lq && col < p.d) v = *(const float4*)(qrow + col);       // d % 4 == 0: a float4 is all inside or all outside
                        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
                        const int chunk = col >> 3, atom = chunk >> 3, cc = chunk & 7;
                        uint2 val;
                        val.x = *(const uint32_t*)&h0;
                        val.y = *(const uint32_t*)&h1;
                        *(uint2*)(sQ + atom * (BLOCK_M * 128) + row * 128 + ((cc ^ (row & 7)) << 4) + ((col & 4) << 1)) = val;
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&q_ready);
        } else {
            const bool valid = qi < p.Lq;
            const float* qrow = (const float*)((const char*)p.q + (int64_t)qi * p.q_nb1 + (int64_t)h * p.q_nb2 + (int64_t)nb * p.q_nb3);
#pragma unroll 1
            for (int c = 0; c < NATOM * 8; ++c) {      // 16-byte chunks of 8 halves
                __half2 hv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = c * 8 + e * 2;
                    float a = (valid && col < p.d) ? qrow[col] : 0.f;
                    float b = (valid && col + 1 < p.d) ? qrow[col + 1] : 0.f;
                    hv[e] = __floats2half2_rn(a, b);
                }
                const int atom = c >> 3, cc = c & 7;
                uint4 val;
                memcpy(&val, hv, 16);
                *(uint4*)(sQ + atom * (BLOCK_M * 128) + r * 128 + ((cc ^ (r & 7)) << 4)) = val;
            }
            fence_proxy_async();
            mbar_arrive(&q_ready);
        }

        float m_ref = -INFINITY, l = 0.f;
        const __half* mrow = nullptr;
        if (p.mask && qi < p.Lq)
            mrow = (const __half*)((const char*)p.mask + (int64_t)qi * p.m_nb1 + (int64_t)(h % p.m_ne2) * p.m_nb2 + (int64_t)(nb % p.m_ne3) * p.m_nb3);

        for (int j = 0; j < nblk; ++j) {
            const int s = j & 1;
            const uint32_t tS = tmem_S0 + s * BLOCK_N + lane_off;
            const int kbase = j * BLOCK_N;
            mbar_wait(&s_full[s], (j >> 1) & 1);
            tc_fence_after();
            // ---- S row: ONE TMEM read into registers, row maximum.  Common case (no mask, all keys of the tile valid): the raw scores stay
            //      in registers and the scale is folded into the exponent's FMA below; otherwise scale, mask and the -inf tail are applied here
            float sv[BLOCK_N];
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            const bool plain = p.pos_scale && (mrow == nullptr) && (kbase + BLOCK_N <= p.Lk);
            const float mul = plain ? p.scale_log2 : 1.0f;
#pragma unroll
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tS + c0, v);
                tmem_ld_wait();
                if (plain) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float t = __uint_as_float(v[i]);
                        sv[c0 + i] = t;
                        mx4[i & 3] = fmaxf(mx4[i & 3], t);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int key = kbase + c0 + i;
                        float t = __uint_as_float(v[i]) * p.scale_log2;
                        if (mrow && key < p.Lk) t += __half2float(mrow[key]) * p.log2e;
                        t = key < p.Lk ? t : -INFINITY;
                        sv[c0 + i] = t;
                        mx4[i & 3] = fmaxf(mx4[i & 3], t);
                    }
                }
            }
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            if (plain) mx *= p.scale_log2;                       // scale > 0: the maximum commutes with it
            const float m_new = fmaxf(m_ref, mx);
            // A row whose keys so far are ALL masked keeps m_ref = -inf (its probabilities are exact zeros); the first finite maximum then
            // forces a rescale with alpha = exp2(-inf) = 0.  (Forcing the reference to 0 here would exponentiate later, much smaller scores
            // against 0 and flush them to f16 zeros: rows of a left-padded mask came out as zeros.)
            bool grow = m_new > -INFINITY && (m_ref == -INFINITY || m_new - m_ref > kLazyThreshold);
            // previous P.V must have retired before P is overwritten or O is rescaled; the (rare) rescale needs it now,
            // otherwise the wait is deferred until the new probabilities sit in registers so the exps overlap that MMA
            bool waited = j == 0;
            if (j > 0 && __any_sync(0xffffffffu, grow)) {
                mbar_wait(&pv_done, (j - 1) & 1);
                tc_fence_after();
                waited = true;
                const float alpha = grow ? fast_exp2(m_ref - m_new) : 1.0f;
                l *= alpha;
#pragma unroll 1
                for (int c0 = 0; c0 < p.dv16; c0 += 16) {
                    uint32_t o[16];
                    tmem_ld16(tmem_O + lane_off + c0, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st16(tmem_O + lane_off + c0, o);
                }
                tmem_st_wait();
            }
            if (grow) m_ref = m_new;
            // ---- pass 2: P = exp2(s - m_ref), row sum, f16 into swizzled shared memory
            float ls4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t ph[BLOCK_N / 2];   // the whole P row as packed half2, kept in registers until the P buffer is free
            const float neg_m = m_ref == -INFINITY ? 0.f : -m_ref;       // all keys masked so far: exp2(-inf + 0) = 0
#pragma unroll
            for (int i = 0; i < BLOCK_N; i += 2) {
                // exp2(s * scale - m): one FMA feeding the MUFU (mul == 1 when the scale was applied above); exp2(-inf) == 0 for masked / tail keys
                const float e0 = fast_exp2(fmaf(sv[i], mul, neg_m)), e1 = fast_exp2(fmaf(sv[i + 1], mul, neg_m));
                const __half2 hv = __floats2half2_rn(e0, e1);
                // accumulate the sum from the ROUNDED probabilities: what the tensor core multiplies is what we normalise by
                const float2 f = __half22float2(hv);
                ls4[(i >> 1) & 3] += f.x + f.y;
                ph[i >> 1] = *(const uint32_t*)&hv;
            }
            l += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
            if (!waited) {
                mbar_wait(&pv_done, (j - 1) & 1);
                tc_fence_after();
            }
#pragma unroll
            for (int chunk = 0; chunk < BLOCK_N / 8; ++chunk) {       // 16-byte chunks of 8 keys
                const int atom = chunk >> 3, cc = chunk & 7;
                const uint4 val = make_uint4(ph[chunk * 4], ph[chunk * 4 + 1], ph[chunk * 4 + 2], ph[chunk * 4 + 3]);
                *(uint4*)(sP + atom * (BLOCK_M * 128) + r * 128 + ((cc ^ (r & 7)) << 4)) = val;
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&p_full);
        }
        // ---- epilogue: O / l -> dst[dv, h, q, n]
        mbar_wait(&pv_done, (nblk - 1) & 1);
        tc_fence_after();
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        float* drow = (float*)((char*)p.dst + (int64_t)h * p.dst_nb1 + (int64_t)qi * p.dst_nb2 + (int64_t)nb * p.dst_nb3);
        __half* drow16 = p.dst16 ? (__half*)((char*)p.dst16 + (((int64_t)h * p.dst_nb1 + (int64_t)qi * p.dst_nb2 + (int64_t)nb * p.dst_nb3) >> 1)) : nullptr;
        const int dv = p.d;
        // f32 result: each warp transposes its 32 rows x 32 columns through shared memory (the K / V rings are dead: every MMA has retired,
        // no TMA is in flight) so that a store instruction writes 128 contiguous bytes of ONE query row instead of 4 bytes of 32 rows
        float* tile = (float*)sK + qd * (32 * 33);
        const char* dbase = (const char*)p.dst + (int64_t)h * p.dst_nb1 + (int64_t)nb * p.dst_nb3;
#pragma unroll 1
        for (int c0 = 0; c0 < p.dv16; c0 += 32) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_off + c0, o);
            tmem_ld_wait();
            if (!p.skip_f32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) tile[lane * 33 + i] = __uint_as_float(o[i]) * inv;
            }
            if (drow16 && qi < p.Lq) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int cb = c0 + hh * 16;
                    if (cb + 16 <= dv) {       // d % 8 == 0 and 16-byte aligned rows: two 16-byte stores
                        uint32_t hp[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __half2 v = __floats2half2_rn(__uint_as_float(o[hh * 16 + 2 * i]) * inv, __uint_as_float(o[hh * 16 + 2 * i + 1]) * inv);
                            hp[i] = *(const uint32_t*)&v;
                        }
                        *(uint4*)(drow16 + cb) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                        *(uint4*)(drow16 + cb + 8) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (cb + i < dv) drow16[cb + i] = __float2half_rn(__uint_as_float(o[hh * 16 + i]) * inv);
                    }
                }
            }
            if (p.skip_f32) continue;
            __syncwarp();
            if (c0 + lane < dv) {
#pragma unroll 4
                for (int rr = 0; rr < 32; ++rr) {
                    const int qrow_i = q0 + qd * 32 + rr;
                    if (qrow_i < p.Lq) *(float*)(dbase + (int64_t)qrow_i * p.dst_nb2 + (int64_t)(c0 + lane) * 4) = tile[rr * 33 + lane];
                }
            }
            __syncwarp();
        }
        (void)drow;
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

bool encode_map(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t s1, uint64_t s2, uint64_t s3,
                uint32_t box0, uint32_t box1) {
    auto enc = b200_get_tensormap_encoder();
    if (!enc) return false;
    cuuint64_t dims[4] = {d0, d1, d2, d3};
    cuuint64_t strides[3] = {s1, s2, s3};
    if (d2 == 1 || strides[1] == 0) strides[1] = strides[0] * dims[1];
    if (d3 == 1 || strides[2] == 0) strides[2] = strides[1] * dims[2];
    cuuint32_t box[4] = {box0, box1, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NATOM, int BLOCK_N>
int launch_fa(cudaStream_t s, dim3 grid, const CUtensorMap& tk, const CUtensorMap& tv, const FaParams& p) {
    using C = FaCfg<NATOM, BLOCK_N>;
    static bool configured[B200_MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev]) {
        if (cudaFuncSetAttribute(k_flash_attn<NATOM, BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) return -1;
        configured[dev] = true;
    }
    b200_launch(k_flash_attn<NATOM, BLOCK_N>, dim3(grid), dim3(192), C::SMEM, s, tk, tv, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace

// vt: packed V^T, f16 [Lk_pad, dv, Hkv, N] dense (row stride Lk_pad elements)
int b200_launch_flash_attn_fused(cudaStream_t s, const b200_td& q, const b200_td& k, const void* vt, int64_t Lk_pad, const b200_td& v,
                                 const b200_td* mask, const b200_td& dst, float scale, void* dst16, int skip_f32) {
    const int64_t d = q.ne[0], Lq = q.ne[1], H = q.ne[2], NB = q.ne[3];
    const int64_t Lk = k.ne[1], Hkv = k.ne[2], dv = v.ne[0];
    if (d != dv || d % 8 || d > 192 || k.type != GGML_TYPE_F16) return -1;
    if (q.nb[0] != 4 || k.nb[0] != 2) return -1;
    if (((uintptr_t)k.data & 15) || (k.nb[1] % 16) || (Hkv > 1 && k.nb[2] % 16) || (NB > 1 && k.nb[3] % 16)) return -1;
    const bool v_mn = vt == nullptr;          // no pre-transposed copy given: read V in place (f16, unit stride along dv, 16-byte aligned strides)
    if (v_mn && (v.type != GGML_TYPE_F16 || v.nb[0] != 2 || ((uintptr_t)v.data & 15) || (v.nb[1] % 16) || (Hkv > 1 && v.nb[2] % 16) || (NB > 1 && v.nb[3] % 16))) return -1;
    if ((!v_mn && ((uintptr_t)vt & 15)) || ((uintptr_t)q.data & 3)) return -1;
    if (Lq == 0 || Lk == 0) return -1;
    const int natom = (int)((d + 63) / 64);
    // d <= 64: 64-key tiles keep a CTA at 65 KB of shared memory and 256 TMEM columns, so two CTAs share an SM and one's
    // softmax overlaps the other's MMAs (the per-tile dependency chain S -> softmax -> P -> PV is latency bound)
    // d in (64, 128]: 64-key tiles as well -- 112 KB of shared memory and 256 TMEM columns per CTA, so two CTAs (two softmax warpgroups,
    // two MMA streams) share an SM; the 128-key variant (one CTA per SM) is kept behind GGML_B200_FA_BN128=1 for A/B runs
    static int bn128 = -1;
    if (bn128 < 0) { const char* e = getenv("GGML_B200_FA_BN128"); bn128 = (e && *e) ? atoi(e) : 0; }
    const int block_n = (natom == 2 && bn128) ? 128 : 64;
    const int dv16 = (int)((dv + 15) / 16 * 16);

    CUtensorMap tk, tv;
    if (!encode_map(&tk, k.data, (uint64_t)d, (uint64_t)Lk, (uint64_t)Hkv, (uint64_t)NB, (uint64_t)k.nb[1], (uint64_t)k.nb[2], (uint64_t)k.nb[3], 64,
                    (uint32_t)block_n))
        return -1;
    if (v_mn) {
        if (!encode_map(&tv, v.data, (uint64_t)dv, (uint64_t)Lk, (uint64_t)Hkv, (uint64_t)NB, (uint64_t)v.nb[1], (uint64_t)v.nb[2], (uint64_t)v.nb[3], 64,
                        (uint32_t)block_n))
            return -1;
    } else if (!encode_map(&tv, vt, (uint64_t)Lk, (uint64_t)dv, (uint64_t)Hkv, (uint64_t)NB, (uint64_t)(Lk_pad * 2), (uint64_t)(Lk_pad * dv * 2),
                           (uint64_t)(Lk_pad * dv * Hkv * 2), 64, (uint32_t)dv16))
        return -1;

    FaParams p;
    memset(&p, 0, sizeof(p));
    p.q = (const float*)q.data;
    p.q_nb1 = q.nb[1]; p.q_nb2 = q.nb[2]; p.q_nb3 = q.nb[3];
    p.dst = (float*)dst.data;
    p.dst_nb1 = dst.nb[1]; p.dst_nb2 = dst.nb[2]; p.dst_nb3 = dst.nb[3];
    // the f16 copy needs 16-byte aligned rows for its vector stores
    p.dst16 = (dst16 && !((uintptr_t)dst16 & 15) && !(dst.nb[1] & 31) && !(dst.nb[2] & 31) && !(dst.nb[3] & 31)) ? (__half*)dst16 : nullptr;
    if (dst16 && !p.dst16) return -1;
    if (skip_f32 && !p.dst16) return -1;
    p.skip_f32 = skip_f32 ? 1 : 0;
    if (mask) {
        p.mask = (const __half*)mask->data;
        p.m_nb1 = mask->nb[1]; p.m_nb2 = mask->nb[2]; p.m_nb3 = mask->nb[3];
        p.m_ne2 = (int)mask->ne[2]; p.m_ne3 = (int)mask->ne[3];
    } else {
        p.m_ne2 = p.m_ne3 = 1;
    }
    p.d = (int)d; p.dv16 = dv16; p.Lq = (int)Lq; p.Lk = (int)Lk; p.H = (int)H; p.rk = (int)(H / Hkv);
    p.v_mn = v_mn ? 1 : 0;
    p.q_vec = (((uintptr_t)q.data & 15) == 0 && (q.nb[1] & 15) == 0 && (q.nb[2] & 15) == 0 && (q.nb[3] & 15) == 0 && (d & 3) == 0) ? 1 : 0;
    p.pos_scale = scale > 0.f ? 1 : 0;
    p.log2e = 1.4426950408889634f;
    p.scale_log2 = scale * p.log2e;
    dim3 grid((unsigned)((Lq + BLOCK_M - 1) / BLOCK_M), (unsigned)H, (unsigned)NB);
    if (H > 65535 || NB > 65535) return -1;
    if (natom == 1) return launch_fa<1, 64>(s, grid, tk, tv, p);
    if (natom == 2) return block_n == 128 ? launch_fa<2, 128>(s, grid, tk, tv, p) : launch_fa<2, 64>(s, grid, tk, tv, p);
    return launch_fa<3, 64>(s, grid, tk, tv, p);
}
