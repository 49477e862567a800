// gemm_tc2.cu -- the CTA-PAIR tcgen05 GEMM / implicit-GEMM convolution: D[n][m] = sum_k A[m][k] * B[n][k]  (F16 / BF16, f32 accumulate)
//
// Why a second kernel.  The one-CTA kernel (gemm_tc.cu) stages (128 + BN) x 128 B per k-block for 128 x BN x 64 MACs; on B200 the
// L2 -> SM path (about 6300 B/clk for the whole chip, about 43 B/clk per SM when all 148 pull -- B300_MICROARCH "LTS cap") caps it
// near 50 % of the tensor pipe and the profile of round 1 showed exactly that (profiles/r01_gemm_notes.md).  Here two CTAs of a
// cluster (one TPC) execute ONE `tcgen05.mma.cta_group::2` of M = 256: each CTA stages its own 128 rows of A and only HALF of the B
// tile, the tensor cores read both halves across the pair, so the bytes per MAC through L2 drop by a third (BN = 256: 85 -> 128
// flop/B) and every B byte is fetched once per 256 output rows.
//
//   * persistent: grid = number of CTA pairs that fit (<= SMs / 2); each pair walks tiles t = pair, pair + P, ...
//   * TMEM holds TWO accumulators (2 x BN columns): the MMA issuer starts the main loop of tile i + 1 while the eight epilogue warps
//     of both CTAs drain tile i (acc_full / acc_empty mbarriers) -- the epilogue, 43 % of the warp samples in round 1, leaves the
//     critical path whenever a pair owns more than one tile
//   * warp roles per CTA (352 threads): warps 0 and 10 = TMA producers for A and B (each CTA loads its A rows and its half of B; completion
//     is signalled on the LEADER CTA's full barrier: `cp.async.bulk.tensor...cta_group::2`), warp 1 = TMEM allocation (both CTAs) and, in the leader
//     only, the single-thread MMA issue + `tcgen05.commit...multicast::cluster` that frees the ring slot in both CTAs,
//     warps 2..9 = epilogue (two warps per TMEM lane quadrant, alternating 32-column chunks)
//   * small-M / small-N problems with a long K: split-K inside the cluster (2, 1, splits): partial tiles stay in shared memory and
//     are reduced through DSMEM in split order (deterministic), as in gemm_tc.cu
//   * convolution mode: A is the NHWC f16 image read as halo boxes {64 ch, BW, BH} (zero fill = the conv's padding), B the packed
//     filter; k-block = (tap, 64-channel block)
//   * convolution with HALO REUSE (3x3, stride 1, pad 1; W % 8 == 0, H % 16 == 0): the per-tap mode above re-stages the image tile nine
//     times, and the main loop is bound by exactly those bytes (43 B/clk per SM).  Here a CTA's 128 output pixels are a 16 x 8 patch;
//     ONE box {64 ch, 10, 16 | 18} lands in shared memory as rows of 128 bytes (pixel-major, 128B swizzle) and the MMAs of tap (kh, kw)
//     read it through a descriptor that starts kw (+ 10 kh) rows into the box with a stride of 10 rows between 8-row groups -- the swizzle
//     is a function of the shared-memory address, so any 128-byte row may start a group (measured: tools/desc_probe.cu, all 56 variants
//     exact).  A ring stage holds the box and the filter tiles of its 3 (one kh) or 9 taps; A bytes per k-block drop from 16 KB to
//     6.7 / 2.5 KB and the loop becomes MMA bound for tile N >= 128
//
// Roofline: tensor pipe.  2 * M * N * K flop per launch; algorithmic bytes (M + N) * K * 2 + M * N * 4.
#include "../b200_ops.h"
#include "b200_launch.cuh"
#include "sm100_ptx.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstring>

using namespace sm100;

namespace {

constexpr int BM = 128;                 // rows of A per CTA (TMEM lanes); the pair computes 256
constexpr int BK_BYTES = 128;
constexpr int A_STAGE_BYTES = BM * BK_BYTES;
constexpr int MAX_STAGES = 10;
constexpr int NTHREADS = 352;             // warps: 0 = A producer, 1 = TMEM / MMA issuer, 2..9 = epilogue, 10 = B producer

struct G2Params {
    float* D;
    int64_t ldd, d_batch_stride;
    int64_t M, N;
    int num_k_blocks, splits;
    int bn;                 // tile N of the pair (multiple of 16, <= 256); each CTA stages bn / 2 rows of B
    int stages, stage_bytes;
    int tiles_m, tiles_n, total_tiles;   // tiles_m counts 256-row pair tiles
    int acc_stride, tmem_cols;
    int ne12, r2;
    const float* bias;
    int bias_mode;          // 0 none, 1 per m, 2 per n
    const float* residual;
    int64_t ldr, r_batch_stride;
    const float* gate;      // per-m factor applied after the activation, before the residual (rounded product, then rounded sum)
    int act;
    int tma_store;          // staged epilogue whose 32-column x 128-row chunks leave through cp.async.bulk.tensor stores (no residual / peer copy)
    int vec_epi;            // staged epilogue: the 4 warps of a column group transpose 32 columns x 128 rows through shared memory and store 512 B per column
    int stage_off;          // byte offset of the two 16 KB staging tiles behind the operand ring
    int nprod;              // TMA producer threads per CTA: 2 (A and B issued by different warps, default) or 1 (A/B of GGML_B200_GEMM2_NPROD)
    int conv, conv_W, conv_KW, conv_cblocks, conv_pad, conv_dil;
    int halo_taps;          // 0: per-tap boxes; 3 | 9: halo reuse, taps per ring stage (num_k_blocks then counts STAGES: 3 * cblocks | cblocks)
    int halo_a_bytes;       // bytes of the image box region at the head of a stage (1024-byte multiple); filter tiles follow
    void* D16;                // optional 16-bit copy of the result (operand of the next contraction), same element layout as D
    int d16_bf16, skip_f32;
    const char* wpf;          // weight operand to request from L2 up front (null: off): row-major [wpf_rows][wpf_kbytes], stride wpf_ld bytes
    int64_t wpf_ld, wpf_rows, wpf_kbytes;
    int wpf_is_a;             // the weights are the A operand (Linear) / the B operand (conv filter)
    float* D2;                // optional mirror of D in the PEER GPU's memory (NVLink mapping, kernels/peer.cu); slot chosen by *d2_seq
    const unsigned* d2_seq;
    int64_t d2_slot;
    // GEGLU mode (> 0: the width `inner` of the gated unit): A holds 2 * inner rows, x-half features [0, inner) then gate-half
    // [inner, 2 inner).  A CTA stages 64 x rows and the 64 gate rows of the SAME features as one 128-row A tile, so TMEM lanes 0..63 / 64..127
    // hold x / gate of features j0 .. j0 + 63 and the epilogue emits x * gelu(gate) as the 16-bit operand [token][inner] of the next
    // Linear -- the f32 projection [2 inner, tokens], the CONT + GELU + MUL passes over it and the operand pack never exist.
    int64_t geglu;
    int res_pf;               // request the residual tile from L2 while the epilogue warps wait for the accumulator (GGML_B200_RES_PREFETCH, default off)
};

// ---- cta_group::2 flavours of the primitives in sm100_ptx.cuh
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once) on the barrier at this shared-memory offset in every CTA of `mask` when all tcgen05 ops issued so far have retired
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// TMA tile load whose completion bytes are counted on a barrier that may live in the peer CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_4d_2cta(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(smem)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

__device__ __forceinline__ float act_fn(float v, int act) {
    if (act == 1) return v / (1.0f + expf(-v));
    if (act == 2) return 0.5f * v * (1.0f + tanhf(0.79788456080286535587989211986876f * v * (1.0f + 0.044715f * v * v)));
    return v;
}

// FMT: 0 = f16, 1 = bf16
template <int FMT>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                         const __grid_constant__ CUtensorMap tmD, const G2Params p) {
    constexpr int BK = 64, UMMA_K = 16;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();          // cluster (2, 1, splits): rank = pair rank + 2 * split
    const uint32_t prank = crank & 1u;                 // 0: leader of the pair (issues the MMAs, owns the full / acc_empty barriers)
    const uint32_t leader = crank & ~1u;
    const int split = (int)(crank >> 1);
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int kb0 = (int)(((int64_t)split * p.num_k_blocks) / p.splits);
    const int kb1 = (int)(((int64_t)(split + 1) * p.num_k_blocks) / p.splits);
    const int nkb = kb1 - kb0;
    const int half_bn = p.bn >> 1;
    const uint32_t my_stage_bytes = (uint32_t)(A_STAGE_BYTES + half_bn * BK_BYTES);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 2 * p.nprod);      // one arrive.expect_tx from each of the two producer threads (A, B) of each CTA of the pair (leader's barrier)
            mbar_init(&empty_bar[s], 1);     // one multicast tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);      // multicast tcgen05.commit
            mbar_init(&acc_empty[b], 16);    // eight epilogue warps of each CTA (used in the leader only)
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        tmem_alloc2(&tmem_base_smem, (uint32_t)p.tmem_cols);
        tmem_relinquish2();
    }
    tc_fence_before();
    cluster_sync_all();           // barriers of every CTA of the cluster are initialised before anyone signals them remotely
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    // tile -> (pair m tile, n tile, batch): m fastest, so the pairs running at one time share the B tile in L2
    auto decode = [&](int t, int& m0, int& n0, int& batch) {
        const int mt = t % p.tiles_m;
        const int r = t / p.tiles_m;
        const int nt = r % p.tiles_n;
        batch = r / p.tiles_n;
        m0 = p.geglu ? mt * BM + (int)prank * (BM / 2) : mt * (2 * BM) + (int)prank * BM;      // this CTA's 128 rows (GEGLU: its 64 features)
        n0 = nt * p.bn;                            // first column of the pair's tile
    };
    if (p.wpf && pair < p.total_tiles) {
        // constant weights: the slab this CTA stages for its first tile (its rows, its K range) is requested from L2 now, before the
        // predecessor kernel has finished -- one bulk prefetch per row segment, spread over the CTA's threads
        int m0, n0, batch;
        decode(pair, m0, n0, batch);
        const int row0 = p.wpf_is_a ? m0 : n0 + (int)prank * half_bn;
        const int nrows = p.wpf_is_a ? BM : half_bn;
        const int64_t koff = (int64_t)kb0 * BK_BYTES;
        const int64_t kbytes = min((int64_t)nkb * BK_BYTES, p.wpf_kbytes - koff) & ~(int64_t)15;
        if (kbytes > 0)
            for (int r = threadIdx.x; r < nrows; r += NTHREADS)
                if (row0 + r < p.wpf_rows) l2_prefetch_bulk(p.wpf + (int64_t)(row0 + r) * p.wpf_ld + koff, (unsigned)kbytes);
    }
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0 || warp == 10) {
        if (lane == 0 && (warp == 0 || p.nprod == 2)) {
            // ===================== TMA producers (both CTAs): warp 0 streams the A tiles, warp 10 the B half-tiles =====================
            // One thread issuing both loads of a k-block was the main-loop limiter on small tiles (about 650 clk per k-block whatever the
            // box size: the issue path of cp.async.bulk.tensor, not bandwidth); two independent issuers halve it.
            const bool is_a = warp == 0;
            const uint32_t full0 = dsmem_map(smem_u32(&full_bar[0]), leader);
            const bool both = p.nprod == 1;        // single-producer variant: this thread issues the B load as well
            const uint32_t my_bytes = both ? my_stage_bytes : (is_a ? (uint32_t)A_STAGE_BYTES : (uint32_t)(half_bn * BK_BYTES));
            uint32_t it = 0;
            for (int t = pair; t < p.total_tiles; t += npairs) {
                int m0, n0, batch;
                decode(t, m0, n0, batch);
                const int i2 = batch % p.ne12, i3 = batch / p.ne12;
                const int nb = n0 + (int)prank * half_bn;          // this CTA's half of the B tile
                int y0 = p.conv ? m0 / p.conv_W : 0, x0 = p.conv ? m0 - y0 * p.conv_W : 0;
                if (p.halo_taps) {
                    // 16 x 8 pixel patches in row-major patch order: patch index = m0 / 128
                    const int tw = p.conv_W >> 3, ti = m0 >> 7;
                    y0 = (ti / tw) * 16; x0 = (ti % tw) * 8;
                    const int T = p.halo_taps;
                    const uint32_t a_bytes = (uint32_t)(10 * (T == 9 ? 18 : 16) * 128), b_bytes = (uint32_t)(T * half_bn * BK_BYTES);
                    const uint32_t mine = both ? a_bytes + b_bytes : (is_a ? a_bytes : b_bytes);
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = (int)(it % (uint32_t)p.stages);
                        const uint32_t ph = (it / (uint32_t)p.stages) & 1u;
                        mbar_wait(&empty_bar[s], ph ^ 1u);
                        const uint32_t fb = full0 + 8u * (uint32_t)s;
                        mbar_expect_tx_cluster(fb, mine);
                        uint8_t* sa = smem + (size_t)s * p.stage_bytes;
                        const int kh = T == 9 ? 0 : kb / p.conv_cblocks, cb = T == 9 ? kb : kb - kh * p.conv_cblocks;
                        if (is_a) tma_load_4d_2cta(sa, &tmA, fb, cb * 64, x0 - 1, y0 + (T == 9 ? 0 : kh) - 1, i2);
                        if (!is_a || both) {
                            for (int tp = 0; tp < T; ++tp) {
                                const int tap = T == 9 ? tp : kh * 3 + tp;
                                tma_load_4d_2cta(sa + p.halo_a_bytes + (size_t)tp * half_bn * BK_BYTES, &tmB, fb, (tap * p.conv_cblocks + cb) * BK, nb, 0, 0);
                            }
                        }
                    }
                    continue;
                }
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = (int)(it % (uint32_t)p.stages);
                    const uint32_t ph = (it / (uint32_t)p.stages) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    const uint32_t fb = full0 + 8u * (uint32_t)s;
                    mbar_expect_tx_cluster(fb, my_bytes);
                    uint8_t* sa = smem + (size_t)s * p.stage_bytes;
                    if (!is_a || both) tma_load_4d_2cta(sa + A_STAGE_BYTES, &tmB, fb, kb * BK, nb, p.conv ? 0 : i2, p.conv ? 0 : i3);
                    if (!is_a) {
                    } else if (p.conv) {
                        const int tap = kb / p.conv_cblocks, cb = kb - tap * p.conv_cblocks;
                        const int kh = tap / p.conv_KW, kw = tap - kh * p.conv_KW;
                        tma_load_4d_2cta(sa, &tmA, fb, cb * 64, x0 + kw * p.conv_dil - p.conv_pad, y0 + kh * p.conv_dil - p.conv_pad, i2);
                    } else if (p.geglu) {
                        // two 64-row boxes: the x rows of this CTA's features, then the gate rows of the same features
                        tma_load_4d_2cta(sa, &tmA, fb, kb * BK, m0, i2 / p.r2, i3);
                        tma_load_4d_2cta(sa + (BM / 2) * BK_BYTES, &tmA, fb, kb * BK, (int)p.geglu + m0, i2 / p.r2, i3);
                    } else {
                        tma_load_4d_2cta(sa, &tmA, fb, kb * BK, m0, i2 / p.r2, i3);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (prank == 0) {
            // ===================== MMA issuer (leader CTA only) =====================
            const uint32_t idesc = make_idesc((uint32_t)FMT, 2 * BM, (uint32_t)p.bn);
            const uint16_t pair_mask = (uint16_t)(3u << leader);
            uint32_t it = 0, j = 0;
            for (int t = pair; t < p.total_tiles; t += npairs, ++j) {
                const uint32_t buf = j & 1u, aph = (j >> 1) & 1u;
                mbar_wait(&acc_empty[buf], aph ^ 1u);           // both CTAs have drained this accumulator (first use: free)
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.acc_stride;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = (int)(it % (uint32_t)p.stages);
                    const uint32_t ph = (it / (uint32_t)p.stages) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + (size_t)s * p.stage_bytes);
                        if (p.halo_taps) {
                            // the image box is in shared memory once; tap (kh, kw) = the same pixels kw (+ 10 kh) rows further in
                            const int T = p.halo_taps;
                            for (int tp = 0; tp < T; ++tp) {
                                const uint32_t arow = (uint32_t)((T == 9 ? (tp / 3) * 10 : 0) + tp % 3);
                                const uint64_t da = make_smem_desc_sw128_sbo(sa + arow * 128u, 10u * 128u);
                                const uint64_t db = make_smem_desc_sw128(sa + (uint32_t)p.halo_a_bytes + (uint32_t)(tp * half_bn * BK_BYTES));
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k) mma_f16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || tp > 0 || k > 0) ? 1u : 0u);
                            }
                        } else {
                            const uint64_t da = make_smem_desc_sw128(sa);
                            const uint64_t db = make_smem_desc_sw128(sa + A_STAGE_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) mma_f16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        }
                        mma_commit_mc(&empty_bar[s], pair_mask);                       // ring slot reusable in BOTH CTAs
                        if (kb == nkb - 1) mma_commit_mc(&acc_full[buf], pair_mask);   // accumulator complete (both CTAs' epilogues)
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 10) {
        // ===================== epilogue (warps 2..9 of both CTAs) =====================
        const int q = warp & 3;                       // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;             // the two warps of a quadrant alternate 32-column chunks
        const int ml = q * 32 + lane;
        const uint32_t acc_empty0 = dsmem_map(smem_u32(&acc_empty[0]), leader);
        uint32_t j = 0;
        for (int t = pair; t < p.total_tiles; t += npairs, ++j) {
            int m0, n0, batch;
            decode(t, m0, n0, batch);
            const uint32_t buf = j & 1u, aph = (j >> 1) & 1u;
            const int64_t m = (int64_t)m0 + ml;
            const bool mvalid = m < p.M;
            const float bias_m = (p.bias_mode == 1 && mvalid) ? p.bias[m] : 0.f;
            const float gate_m = (p.gate && mvalid) ? p.gate[m] : 1.f;
            const int ncols = (int)min((int64_t)p.bn, p.N - n0);
            float* Dp = p.D + (int64_t)batch * p.d_batch_stride;
            const float* Rp = p.residual ? p.residual + (int64_t)batch * p.r_batch_stride : nullptr;
            // fused collective: the same element also goes to the peer GPU (offset of D's element inside the mailbox slot)
            const int64_t d2off = p.D2 ? (p.D2 - p.D) + (int64_t)((*p.d2_seq + 1u) & 1u) * p.d2_slot : 0;
            if (Rp && p.res_pf && p.splits == 1 && p.vec_epi) {
                // the main loop of this tile is still running and these eight warps have nothing to do: ask L2 for the residual values the
                // epilogue will add (this CTA's 128 rows of every column of the tile) -- a hint, bounded to rows < M and columns < N
                const int et = (int)threadIdx.x - 64;                  // 0 .. 255
                if (p.halo_taps) {
                    const int tw = p.conv_W >> 3, ti = m0 >> 7;
                    const int64_t base_row = (int64_t)((ti / tw) * 16) * p.conv_W + (ti % tw) * 8;
                    if ((int64_t)m0 < p.M)
                        for (int u = et; u < ncols * 16; u += 256) {
                            const int n = u >> 4, py = u & 15;
                            l2_prefetch_bulk(Rp + (int64_t)(n0 + n) * p.ldr + base_row + (int64_t)py * p.conv_W, 32u);
                        }
                } else {
                    const int64_t rows = min((int64_t)BM, p.M - m0);
                    if (rows > 0)
                        for (int n = et; n < ncols; n += 256) l2_prefetch_bulk(Rp + (int64_t)(n0 + n) * p.ldr + m0, (unsigned)(rows * 4));
                }
            }
            mbar_wait(&acc_full[buf], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)p.acc_stride;
            if (p.splits > 1) {
                // partial tile -> own shared memory [bn][128] f32 (the operand ring is dead: every MMA of this CTA pair has retired)
                float* sred = (float*)smem;
#pragma unroll 1
                for (int c0 = half * 32; c0 < p.bn; c0 += 64) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i < p.bn) sred[(c0 + i) * BM + ml] = __uint_as_float(r[i]);
                }
            } else if (p.geglu) {
                // GEGLU: staged like the vectorised epilogue below.  After the transpose through shared memory lane l holds rows 4l .. 4l + 3 of a
                // column (token): lanes 0..15 the x values of features j0 + 4l .., lanes 16..31 the gate values of features j0 + 4 (l - 16) ..:
                // the gate lanes apply bias + GELU and hand their four values to the x lanes 16 below, which multiply and store 8 bytes.
                // Arithmetic is the unfused chain's: (acc + bias) per half, gelu_tanh in f32, one multiply, one rounding to the 16-bit operand.
                float* stage = (float*)(smem + p.stage_off) + half * (32 * BM);
                const int wq = (warp - 2) & 3;
                const bool is_gate = lane >= 16;
                const int64_t feat = (int64_t)m0 + 4 * (lane & 15);
                const bool fvalid = feat < p.geglu;                       // inner % 4 == 0: the four features are valid together
                float4 bm4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias_mode == 1 && fvalid) bm4 = *(const float4*)(p.bias + (is_gate ? p.geglu : 0) + feat);
                uint16_t* out16 = (uint16_t*)p.D16 + (int64_t)batch * p.N * p.geglu + feat;
#pragma unroll 1
                for (int c0 = half * 32; c0 < ncols; c0 += 64) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) stage[i * BM + ml] = __uint_as_float(r[i]);
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int col = wq * 8 + jj, n = c0 + col;
                        float4 v = *(const float4*)(stage + col * BM + 4 * lane);
                        v.x += bm4.x + 0.f; v.y += bm4.y + 0.f; v.z += bm4.z + 0.f; v.w += bm4.w + 0.f;      // (the plain epilogue's acc + (bias_m + bias_n))
                        if (is_gate) { v.x = act_fn(v.x, 2); v.y = act_fn(v.y, 2); v.z = act_fn(v.z, 2); v.w = act_fn(v.w, 2); }
                        const float gx = __shfl_down_sync(0xffffffffu, v.x, 16), gy = __shfl_down_sync(0xffffffffu, v.y, 16);
                        const float gz = __shfl_down_sync(0xffffffffu, v.z, 16), gw = __shfl_down_sync(0xffffffffu, v.w, 16);
                        if (!is_gate && fvalid && n < ncols) {
                            const float ox = __fmul_rn(v.x, gx), oy = __fmul_rn(v.y, gy), oz = __fmul_rn(v.z, gz), ow = __fmul_rn(v.w, gw);
                            uint2 h;
                            if (p.d16_bf16) {
                                const __nv_bfloat162 a = __floats2bfloat162_rn(ox, oy), b = __floats2bfloat162_rn(oz, ow);
                                h.x = *(const uint32_t*)&a; h.y = *(const uint32_t*)&b;
                            } else {
                                const __half2 a = __floats2half2_rn(ox, oy), b = __floats2half2_rn(oz, ow);
                                h.x = *(const uint32_t*)&a; h.y = *(const uint32_t*)&b;
                            }
                            *(uint2*)(out16 + (int64_t)(n0 + n) * p.geglu) = h;
                        }
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                }
            } else if (p.tma_store) {
                // Staged stores through the TMA: the four warps of this column group write bias / activation applied values of a 32-column
                // chunk to shared memory as [column][row] -- exactly the dense box {128 m, 32 n} of the f32 output map -- and ONE thread hands
                // the 16 KB to cp.async.bulk.tensor.  The LSU path (st.global from 8 warps) topped out near 10 B/clk per SM, as long as a
                // 17-k-block main loop per 128 x 128 tile; the bulk store also clips the M / N edges itself.
                float* stage = (float*)(smem + p.stage_off) + half * (32 * BM);
                const int wq = (warp - 2) & 3;
                const float* bias_n = p.bias_mode == 2 ? p.bias + n0 : nullptr;
#pragma unroll 1
                for (int c0 = half * 32; c0 < ncols; c0 += 64) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    const float bn = (bias_n && c0 + lane < ncols) ? bias_n[c0 + lane] : 0.f;   // lane i carries the bias of column c0 + i
                    tmem_ld_wait();
                    // the previous chunk's bulk store has finished READING the staging tile (its issuer waited) before anyone overwrites it
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float v = __uint_as_float(r[i]) + bias_m + __shfl_sync(0xffffffffu, bn, i);
                        if (p.act) v = act_fn(v, p.act);
                        stage[i * BM + ml] = v;
                    }
                    fence_proxy_async();                                   // generic-proxy writes -> visible to the bulk-copy engine
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                    if (wq == 0 && lane == 0) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tmD)),
                                     "r"(smem_u32(stage)), "r"(m0), "r"(n0 + c0), "r"(batch)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                }
            } else if (p.vec_epi) {
                // Staged stores.  A thread owns one accumulator ROW (TMEM lane), but memory wants a column's 128 consecutive rows in one
                // piece: the four warps of this column group write a 32-column chunk to shared memory as [column][row], then every warp
                // streams 8 of the columns with 16 bytes per lane -- one store instruction = 512 contiguous bytes (and the residual is read
                // the same way), instead of 4 x 128 B scattered over four instructions per column.
                float* stage = (float*)(smem + p.stage_off) + half * (32 * BM);
                const int wq = (warp - 2) & 3;
                const int64_t mrow = (int64_t)m0 + 4 * lane;
                const bool rvalid = mrow < p.M;                           // M % 4 == 0 on this path
                // where this lane's four consecutive rows live inside an image plane: m itself, or -- halo mode, 16 x 8 pixel patches --
                // four neighbours of one image row
                int64_t prow = mrow;
                if (p.halo_taps) {
                    const int tw = p.conv_W >> 3, ti = m0 >> 7;
                    prow = (int64_t)((ti / tw) * 16 + (lane >> 1)) * p.conv_W + (ti % tw) * 8 + (lane & 1) * 4;
                }
                float4 bm4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias_mode == 1 && rvalid) bm4 = *(const float4*)(p.bias + mrow);
                float4 gm4 = make_float4(1.f, 1.f, 1.f, 1.f);
                if (p.gate && rvalid) gm4 = *(const float4*)(p.gate + mrow);
#pragma unroll 1
                for (int c0 = half * 32; c0 < ncols; c0 += 64) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) stage[i * BM + ml] = __uint_as_float(r[i]);
                    // The residual values of this warp's eight columns are requested together and BEFORE the barrier (the accumulator registers
                    // are dead by now, so this costs no registers): one memory latency per chunk instead of eight in a row -- D and the residual
                    // may alias (in-place add), which kept the compiler from hoisting the loads itself, and a conv with a residual ran at
                    // 240 us where the same conv without one took 107 us (512 x 512 x 128 VAE level).  A thread reads exactly the elements it
                    // writes, and all eight reads precede the eight writes, so the in-place case stays correct.
                    float4 rr8[8];
                    float bn8[8];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int n = c0 + wq * 8 + jj;
                        rr8[jj] = (Rp && n < ncols && rvalid) ? *(const float4*)(Rp + (int64_t)(n0 + n) * p.ldr + prow) : make_float4(0.f, 0.f, 0.f, 0.f);
                        bn8[jj] = (p.bias_mode == 2 && n < ncols) ? p.bias[n0 + n] : 0.f;
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int col = wq * 8 + jj, n = c0 + col;
                        if (n < ncols && rvalid) {
                            float4 v = *(const float4*)(stage + col * BM + 4 * lane);
                            const float bn = bn8[jj];
                            v.x += bm4.x + bn; v.y += bm4.y + bn; v.z += bm4.z + bn; v.w += bm4.w + bn;
                            if (p.act) { v.x = act_fn(v.x, p.act); v.y = act_fn(v.y, p.act); v.z = act_fn(v.z, p.act); v.w = act_fn(v.w, p.act); }
                            if (p.gate) { v.x = __fmul_rn(v.x, gm4.x); v.y = __fmul_rn(v.y, gm4.y); v.z = __fmul_rn(v.z, gm4.z); v.w = __fmul_rn(v.w, gm4.w); }
                            const int64_t off = (int64_t)(n0 + n) * p.ldd + prow;
                            if (Rp) {
                                const float4 rr = rr8[jj];
                                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
                            }
                            if (!p.skip_f32) *(float4*)(Dp + off) = v;
                            if (p.D2) *(float4*)(Dp + off + d2off) = v;
                            if (p.D16) {
                                uint2 h;
                                if (p.d16_bf16) {
                                    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
                                    h.x = *(const uint32_t*)&a; h.y = *(const uint32_t*)&b;
                                } else {
                                    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
                                    h.x = *(const uint32_t*)&a; h.y = *(const uint32_t*)&b;
                                }
                                *(uint2*)((uint16_t*)p.D16 + (int64_t)batch * p.d_batch_stride + off) = h;
                            }
                        }
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                }
            } else if (p.act == 0) {
                float* dptr = Dp + (int64_t)n0 * p.ldd + m;
                const float* rptr = Rp ? Rp + (int64_t)n0 * p.ldr + m : nullptr;
                const float* bias_n = p.bias_mode == 2 ? p.bias + n0 : nullptr;
#pragma unroll 1
                for (int c0 = half * 32; c0 < ncols; c0 += 64) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    const float bn = (bias_n && c0 + lane < ncols) ? bias_n[c0 + lane] : 0.f;   // lane i carries the bias of column c0 + i
                    tmem_ld_wait();
                    if (rptr == nullptr) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float v = __uint_as_float(r[i]) + bias_m + __shfl_sync(0xffffffffu, bn, i);
                            if (mvalid && c0 + i < ncols) {
                                dptr[(int64_t)(c0 + i) * p.ldd] = v;
                                if (p.D2) dptr[(int64_t)(c0 + i) * p.ldd + d2off] = v;
                            }
                        }
                    } else {
                        float rr[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) rr[i] = (mvalid && c0 + i < ncols) ? rptr[(int64_t)(c0 + i) * p.ldr] : 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float v = __uint_as_float(r[i]) + bias_m + __shfl_sync(0xffffffffu, bn, i);
                            if (p.gate) v = __fmul_rn(v, gate_m);
                            v += rr[i];
                            if (mvalid && c0 + i < ncols) {
                                dptr[(int64_t)(c0 + i) * p.ldd] = v;
                                if (p.D2) dptr[(int64_t)(c0 + i) * p.ldd + d2off] = v;
                            }
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c0 = half * 32; c0 < ncols; c0 += 64) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int64_t n = (int64_t)n0 + c0 + i;
                        if (mvalid && c0 + i < ncols) {
                            float v = __uint_as_float(r[i]) + bias_m;
                            if (p.bias_mode == 2) v += p.bias[n];
                            v = act_fn(v, p.act);
                            if (p.gate) v = __fmul_rn(v, gate_m);
                            if (Rp) v += Rp[n * p.ldr + m];
                            Dp[n * p.ldd + m] = v;
                            if (p.D2) Dp[n * p.ldd + m + d2off] = v;
                        }
                    }
                }
            }
            // every tcgen05.ld of this accumulator has completed: hand it back to the MMA issuer of the pair
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty0 + 8u * buf);
        }
    }

    if (p.splits > 1) {
        // ---- split-K reduction across the cluster: CTA (prank, split) owns the columns [split * bn / S, (split + 1) * bn / S) of its
        //      128 rows and sums the partial tiles of the CTAs with the same pair rank in split order (deterministic)
        tc_fence_before();
        cluster_sync_all();
        if (warp >= 2 && warp < 10) {
            int m0, n0, batch;
            decode(pair, m0, n0, batch);
            const int w8 = warp - 2;
            const int64_t mrow = (int64_t)m0 + 4 * lane;
            int64_t prow = mrow;                  // offset inside the image plane (see the staged epilogue)
            if (p.halo_taps) {
                const int tw = p.conv_W >> 3, ti = m0 >> 7;
                prow = (int64_t)((ti / tw) * 16 + (lane >> 1)) * p.conv_W + (ti % tw) * 8 + (lane & 1) * 4;
            }
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
            for (int s = 0; s < 8; ++s) peer[s] = s < p.splits ? dsmem_map(sred_local, prank + 2u * (uint32_t)s) : 0u;
            const int ncols = (int)min((int64_t)p.bn, p.N - n0);
            const int cbeg = (split * p.bn) / p.splits, cend = min(((split + 1) * p.bn) / p.splits, ncols);
            const bool vec_ok = (mrow + 3 < p.M) && ((p.ldd & 3) == 0) && ((((uintptr_t)Dp) & 15) == 0) && ((m0 & 3) == 0);
            const int64_t d2off = p.D2 ? (p.D2 - p.D) + (int64_t)((*p.d2_seq + 1u) & 1u) * p.d2_slot : 0;
            // the residual of the NEXT column is requested before the partial tiles of this one are summed (same aliasing argument as the
            // staged epilogue: a thread reads only elements it writes itself, one iteration ahead of the write of a different column)
            const bool res_vec = vec_ok && Rp != nullptr && ((p.ldr & 3) == 0) && ((((uintptr_t)Rp) & 15) == 0);
            float4 rnext = make_float4(0.f, 0.f, 0.f, 0.f);
            if (res_vec && cbeg + w8 < cend) rnext = *(const float4*)(Rp + ((int64_t)n0 + cbeg + w8) * p.ldr + prow);
#pragma unroll 1
            for (int c = cbeg + w8; c < cend; c += 8) {
                const float4 rcur = rnext;
                if (res_vec && c + 8 < cend) rnext = *(const float4*)(Rp + ((int64_t)n0 + c + 8) * p.ldr + prow);
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
                if (p.act) { v.x = act_fn(v.x, p.act); v.y = act_fn(v.y, p.act); v.z = act_fn(v.z, p.act); v.w = act_fn(v.w, p.act); }
                if (p.gate) { v.x = __fmul_rn(v.x, gm.x); v.y = __fmul_rn(v.y, gm.y); v.z = __fmul_rn(v.z, gm.z); v.w = __fmul_rn(v.w, gm.w); }
                float* dst = Dp + n * p.ldd + prow;
                if (vec_ok && (Rp == nullptr || (((p.ldr & 3) == 0) && ((((uintptr_t)Rp) & 15) == 0)))) {
                    if (Rp) { const float4 rr = rcur; v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
                    if (!p.skip_f32) *(float4*)dst = v;
                    if (p.D16) {        // (launcher: 8-byte aligned 16-bit rows whenever D16 is set with splits > 1)
                        uint2 h;
                        if (p.d16_bf16) {
                            const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
                            h.x = *(const uint32_t*)&a; h.y = *(const uint32_t*)&b;
                        } else {
                            const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
                            h.x = *(const uint32_t*)&a; h.y = *(const uint32_t*)&b;
                        }
                        *(uint2*)((uint16_t*)p.D16 + (int64_t)batch * p.d_batch_stride + n * p.ldd + prow) = h;
                    }
                    if (p.D2) {
                        if ((d2off & 3) == 0) *(float4*)(dst + d2off) = v;
                        else { dst[d2off] = v.x; dst[d2off + 1] = v.y; dst[d2off + 2] = v.z; dst[d2off + 3] = v.w; }
                    }
                } else {
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (mrow + u < p.M) {
                            const float o = vv[u] + (Rp ? Rp[n * p.ldr + prow + u] : 0.f);
                            dst[u] = o;
                            if (p.D2) dst[u + d2off] = o;
                        }
                }
            }
        }
    }
    if (p.tma_store && lane == 0 && (warp == 2 || warp == 6)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // bulk stores complete
    // nobody may exit while a peer can still read its shared memory (DSMEM reduce, the pair's MMAs) or signal its barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------ host side
bool encode_rows(CUtensorMap* out, const void* ptr, int type, int64_t K, int64_t rows, int64_t ld_elems, int64_t b2, int64_t b2_stride, uint32_t box_rows) {
    auto enc = b200_get_tensormap_encoder();
    if (!enc) return false;
    CUtensorMapDataType dt = type == GGML_TYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)b2, 1};
    cuuint64_t strides[3] = {(cuuint64_t)(ld_elems * 2), (cuuint64_t)(b2_stride * 2), 0};
    if (b2 == 1 || strides[1] == 0) strides[1] = strides[0] * dims[1];
    strides[2] = strides[1] * dims[2];
    cuuint32_t box[4] = {64, box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(out, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int FMT>
cudaError_t launch2(cudaStream_t s, unsigned ctas, unsigned splits, size_t smem, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td,
                    const G2Params& kp) {
    static size_t configured[B200_MAX_DEVICES] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (configured[d] < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc2<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        if (e != cudaSuccess) return e;
        configured[d] = 227 * 1024;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas, 1, splits);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (b200_pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = splits;
    ++n;
    cfg.numAttrs = n;
    cfg.attrs = attr;
    return cudaLaunchKernelEx(&cfg, k_gemm_tc2<FMT>, ta, tb, td, kp);
}

// may the epilogue use 16-byte stores along m?  (GGML_B200_GEMM2_VEC_EPI=0 keeps the per-row path for A/B runs)
int vec_epilogue_ok(const G2Params& kp, int splits) {
    static int en = -1;
    if (en < 0) { const char* e = getenv("GGML_B200_GEMM2_VEC_EPI"); en = (e && *e) ? atoi(e) : 1; }
    if (!en || splits != 1) return 0;
    if ((kp.M & 3) || (kp.ldd & 3) || (kp.d_batch_stride & 3) || ((uintptr_t)kp.D & 15)) return 0;
    if (kp.residual && ((kp.ldr & 3) || (kp.r_batch_stride & 3) || ((uintptr_t)kp.residual & 15))) return 0;
    if (kp.bias_mode == 1 && ((uintptr_t)kp.bias & 15)) return 0;
    if (kp.gate && ((uintptr_t)kp.gate & 15)) return 0;
    if (kp.D2 && (((uintptr_t)kp.D2 & 15) || (kp.d2_slot & 3))) return 0;
    return 1;
}

// f32 output as a 3-D tensor (m, n, batch) with a dense box {128, 32, 1}: what one bulk store of the staged epilogue writes
bool encode_output(CUtensorMap* out, G2Params& kp, int64_t batch) {
    static int en = -1;
    if (en < 0) { const char* e = getenv("GGML_B200_GEMM2_TMA_STORE"); en = (e && *e) ? atoi(e) : 0; }     // measured 1-5 us SLOWER than the staged st.global path (profiles/r02_gemm_model.md): off
    memset(out, 0, sizeof(*out));
    kp.tma_store = 0;
    if (!en || !kp.vec_epi || kp.residual || kp.D2 || kp.D16 || kp.gate) return true;
    if ((kp.ldd * 4) % 16 || (kp.d_batch_stride * 4) % 16 || ((uintptr_t)kp.D & 15)) return true;
    // batch must be the outermost dimension of the output (the P.V product of the unfused attention interleaves heads INSIDE a row:
    // d_batch_stride < ldd -- such maps are left to the st.global path)
    if (batch > 1 && kp.d_batch_stride < kp.ldd * kp.N) return true;
    auto enc = b200_get_tensormap_encoder();
    if (!enc) return true;
    cuuint64_t dims[3] = {(cuuint64_t)kp.M, (cuuint64_t)kp.N, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)kp.ldd * 4, (cuuint64_t)(batch > 1 ? kp.d_batch_stride : kp.ldd * kp.N) * 4};
    cuuint32_t box[3] = {128, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, kp.D, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return true;
    kp.tma_store = 1;
    return true;
}

// shared by the GEMM and the conv front end: fills the tile geometry for a chosen (bn, splits)
// halo_taps: 0 = one (A tile, B tile) pair per stage; 3 | 9 = halo-reuse convolution, a stage = the image box + that many filter tiles
// (nkb then counts stages)
bool fill_geometry(G2Params& kp, const b200_device_info& dev, int64_t M, int64_t N, int64_t batch, int nkb, int bn, int splits, unsigned* ctas, size_t* smem,
                   int halo_taps = 0) {
    static int res_pf = -1;
    if (res_pf < 0) { const char* e = getenv("GGML_B200_RES_PREFETCH"); res_pf = (e && *e) ? atoi(e) : 0; }     // measured on B200: VAE decode 6.0 ms without, 6.2 ms with the request (2048 32-byte requests per halo tile); SD1.5 neutral: off
    kp.res_pf = res_pf;
    static int nprod = -1;
    if (nprod < 0) { const char* e = getenv("GGML_B200_GEMM2_NPROD"); nprod = (e && *e && atoi(e) == 1) ? 1 : 2; }
    kp.nprod = nprod;
    const int sms = dev.sm_count > 0 ? dev.sm_count : 148;
    kp.bn = bn;
    kp.splits = splits;
    kp.num_k_blocks = nkb;
    kp.halo_taps = halo_taps;
    kp.halo_a_bytes = halo_taps == 9 ? 23 * 1024 : (halo_taps == 3 ? 20 * 1024 : 0);       // 10 x 18 | 10 x 16 rows of 128 B, rounded to 1 KB
    kp.stage_bytes = halo_taps ? kp.halo_a_bytes + halo_taps * (bn / 2) * BK_BYTES : A_STAGE_BYTES + (bn / 2) * BK_BYTES;
    const int stage_tiles = splits == 1 ? 2 * 32 * BM * 4 : 0;        // two 16 KB staging tiles of the vectorised epilogue
    int stages = (int)((227 * 1024 - 4096 - stage_tiles) / kp.stage_bytes);
    stages = std::min(stages, MAX_STAGES);
    if (splits > 1) {
        // the partial tile [bn][128] f32 reuses the ring
        while ((size_t)stages * kp.stage_bytes < (size_t)bn * BM * 4) ++stages;
        if ((size_t)stages * kp.stage_bytes + 2048 > 227 * 1024 - 2048 || stages > MAX_STAGES) return false;
    }
    if (stages < (halo_taps == 9 ? 2 : 3)) return false;
    kp.stages = stages;
    int cols = 32;
    while (cols < 2 * bn) cols <<= 1;
    if (cols > 512) return false;
    kp.tmem_cols = cols;
    kp.acc_stride = cols / 2;
    const int64_t tm = (M + 2 * BM - 1) / (2 * BM), tn = (N + bn - 1) / bn;
    const int64_t total = tm * tn * batch;
    if (total <= 0 || total > 0x3fffffff) return false;
    kp.tiles_m = (int)tm; kp.tiles_n = (int)tn; kp.total_tiles = (int)total;
    int64_t pairs = total;
    if (splits == 1) pairs = std::min<int64_t>(total, sms / 2);
    else if (pairs * 2 * splits > 65535 * 2) return false;
    *ctas = (unsigned)(pairs * 2);
    kp.stage_off = stages * kp.stage_bytes;
    *smem = (size_t)stages * kp.stage_bytes + stage_tiles + 1024;
    return true;
}

}  // namespace

// Modelled cycles of the pair kernel for (bn, splits); used by the plan choosers of gemm_tc.cu to pick between the kernels.
// Calibrated on B200 with tools/gemm_bench (profiles/r02_gemm_model.md).  The quantity that decides everything below the MMA floor is
// the SM's ingest port: one SM takes ~43 B/clk from L2 through TMA whether it runs alone or with 147 others (the "6300 B/clk chip
// cap" is 148 such ports), so a CTA's main loop costs (bytes it stages) / 43 clk and small problems are won by spreading the operand
// bytes over as many SMs as possible -- not by bigger tiles.
double b200_gemm_tc2_model(const b200_device_info& dev, int64_t M, int64_t N, int64_t batch, int nkb, int bn, int splits, double a_bytes) {
    const int sms = dev.sm_count > 0 ? dev.sm_count : 148;
    const int64_t tm = (M + 255) / 256, tn = (N + bn - 1) / bn;
    const int64_t tiles = tm * tn * batch;
    const double kb = (double)((nkb + splits - 1) / splits);
    const double ingest = 43.0;                                            // B/clk per SM
    const double bytes = a_bytes + bn / 2.0 * 128.0;                       // per k-block and CTA: its A rows (16 KB, less with halo reuse) + half the B tile
    const double kb_cycles = std::max(2.0 * bn, bytes / ingest);           // cta_group::2: bn / 2 clk per K = 16 MMA, four per k-block
    const double epi = 18.0 * bn + 600.0;                                  // eight epilogue warps: TMEM -> registers -> coalesced f32 stores
    if (splits == 1) {
        const int64_t pairs = std::min<int64_t>(tiles, sms / 2);
        const double per_pair = (double)((tiles + pairs - 1) / pairs);
        // tiles of one pair overlap their epilogue with the next main loop (two TMEM accumulators); the last epilogue is exposed
        return 5500.0 + per_pair * std::max(kb * kb_cycles, epi) + epi;
    }
    // split-K: every cluster owns one tile; the partial tiles are reduced through DSMEM (~17 B/clk per CTA): bn * 128 * 4 B per CTA
    const int csize = 2 * splits;
    const int64_t max_clusters = csize == 4 ? 36 : (csize == 6 ? 20 : 13);   // concurrent clusters the GPCs (16-20 SMs each) can host
    const double waves = (double)((tiles + max_clusters - 1) / max_clusters);
    return waves * (6500.0 + kb * kb_cycles + 31.0 * bn + 900.0);
}

// returns 1 when launched, -1 when the problem is outside this kernel's envelope.  bn / splits <= 0: choose here.
int b200_launch_gemm_tc2(cudaStream_t s, const b200_device_info& dev, const b200_gemm_args& g, int bn, int splits) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0 || g.K <= 0) return -1;
    if (g.type != GGML_TYPE_F16 && g.type != GGML_TYPE_BF16) return -1;
    if (((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15) || (g.lda * 2) % 16 || (g.ldb * 2) % 16) return -1;
    if ((g.a_batch_stride * 2) % 16 || (g.b_batch_stride * 2) % 16) return -1;
    if (g.gate && !g.residual) return -1;       // the gated epilogue exists on the residual paths only
    if (bn < 16 || bn > 256 || (bn & 15) || splits < 1 || splits > 4) return -1;
    const int nkb = (int)((g.K + 63) / 64);
    if (splits > nkb) return -1;
    G2Params kp;
    memset(&kp, 0, sizeof(kp));
    unsigned ctas = 0;
    size_t smem = 0;
    if (!fill_geometry(kp, dev, g.M, g.N, g.batch, nkb, bn, splits, &ctas, &smem)) return -1;
    CUtensorMap ta, tb;
    const int64_t a_batches = (g.batch + g.a_bcast - 1) / g.a_bcast;
    if (g.geglu) {
        // GEGLU mode: whole 64-feature slabs, the 16-bit operand as the only output, bias per feature or none, no other epilogue work
        if (g.M != 2 * g.geglu || g.geglu % 64 || splits != 1 || !g.D16 || !g.skip_f32 || g.residual || g.gate || g.act || (g.bias && g.bias_mode != 1)) return -1;
        if (((uintptr_t)g.D16 & 15) || (g.bias && ((uintptr_t)g.bias & 15)) || (g.d16_type != GGML_TYPE_F16 && g.d16_type != GGML_TYPE_BF16)) return -1;
    }
    if (!encode_rows(&ta, g.A, g.type, g.K, g.M, g.lda, a_batches, g.a_batch_stride, g.geglu ? BM / 2 : BM)) return -1;
    if (!encode_rows(&tb, g.B, g.type, g.K, g.N, g.ldb, g.batch, g.b_batch_stride, (uint32_t)(bn / 2))) return -1;
    kp.D = g.D; kp.ldd = g.ldd; kp.d_batch_stride = g.d_batch_stride;
    kp.M = g.M; kp.N = g.N;
    kp.ne12 = (int)g.batch; kp.r2 = (int)g.a_bcast;
    kp.bias = g.bias; kp.bias_mode = g.bias ? g.bias_mode : 0;
    kp.residual = g.residual; kp.ldr = g.ldr; kp.r_batch_stride = g.d_batch_stride;
    kp.act = g.act;
    kp.gate = g.gate;
    if ((g.wprefetch & 1) && a_batches == 1) { kp.wpf = (const char*)g.A; kp.wpf_ld = g.lda * 2; kp.wpf_rows = g.M; kp.wpf_kbytes = g.K * 2; kp.wpf_is_a = 1; }
    else if ((g.wprefetch & 2) && g.batch == 1) { kp.wpf = (const char*)g.B; kp.wpf_ld = g.ldb * 2; kp.wpf_rows = g.N; kp.wpf_kbytes = g.K * 2; kp.wpf_is_a = 0; }
    kp.geglu = g.geglu;
    if (g.geglu) kp.wpf = nullptr;            // (the up-front L2 request assumes 128 consecutive A rows per CTA)
    kp.vec_epi = g.geglu ? 0 : vec_epilogue_ok(kp, splits);
    // 16-bit copy: the staged epilogue (splits == 1) or the split-K reduce, both with 8-byte stores of four consecutive rows
    const bool d16_split_ok = splits > 1 && !g.residual && !(kp.M & 3) && !(kp.ldd & 3) && !(kp.d_batch_stride & 3) && !((uintptr_t)kp.D & 15);
    if (g.geglu) {
        kp.D16 = g.D16; kp.d16_bf16 = g.d16_type == GGML_TYPE_BF16; kp.skip_f32 = 1;
    } else if (g.D16 && (kp.vec_epi || d16_split_ok) && !((uintptr_t)g.D16 & 7) && (g.d16_type == GGML_TYPE_F16 || g.d16_type == GGML_TYPE_BF16)) {
        kp.D16 = g.D16; kp.d16_bf16 = g.d16_type == GGML_TYPE_BF16; kp.skip_f32 = g.skip_f32;
    } else if (g.D16 && g.d16_strict) {
        return -1;          // the one-CTA kernel may still take it
    }
    CUtensorMap td;
    encode_output(&td, kp, g.batch);
    cudaError_t e = g.type == GGML_TYPE_F16 ? launch2<0>(s, ctas, (unsigned)splits, smem, ta, tb, td, kp) : launch2<1>(s, ctas, (unsigned)splits, smem, ta, tb, td, kp);
    if (e != cudaSuccess) {
        fprintf(stderr, "[ggml-b200] CTA-pair GEMM launch failed: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
        return -1;
    }
    if (kp.D16 && g.d16_done) *g.d16_done = 1;
    return 1;
}

// Diagnostic C-ABI (include/ggml-b200.h), host only: the launch geometry of the pair kernel for a tile width / split-K factor / taps per
// image box -- ring stages, bytes per stage, dynamic shared memory, TMEM columns, CTAs.  tests/test_cabi.py sweeps every legal plan and
// checks the invariants the kernel relies on (1024-byte aligned swizzle atoms, shared-memory and TMEM budgets, split-K partial tile fits the ring).
extern "C" int ggml_backend_b200_debug_pair_geometry(int64_t M, int64_t N, int64_t batch, int nkb, int bn, int splits, int halo_taps, int sm_count, int* stages,
                                                     int* stage_bytes, int64_t* smem_bytes, int* tmem_cols, int* ctas) {
    if (bn < 16 || bn > 256 || (bn & 15) || splits < 1 || splits > 4 || (halo_taps != 0 && halo_taps != 3 && halo_taps != 9)) return 0;
    b200_device_info dev;
    memset(&dev, 0, sizeof(dev));
    dev.sm_count = sm_count;
    G2Params kp;
    memset(&kp, 0, sizeof(kp));
    unsigned c = 0;
    size_t smem = 0;
    if (!fill_geometry(kp, dev, M, N, batch, nkb, bn, splits, &c, &smem, halo_taps)) return 0;
    if (stages) *stages = kp.stages;
    if (stage_bytes) *stage_bytes = kp.stage_bytes;
    if (smem_bytes) *smem_bytes = (int64_t)smem;
    if (tmem_cols) *tmem_cols = kp.tmem_cols;
    if (ctas) *ctas = (int)c;
    return 1;
}

// taps per ring stage the halo-reuse convolution would use for (bn, splits): 9 (whole 3x3 neighbourhood from one box) when two such
// stages fit the shared memory, else 3 (one filter row per box), 0 when neither fits
int b200_conv_tc2_halo_taps(int bn, int splits) {
    if (bn < 16 || bn > 256 || (bn & 15)) return 0;
    const int stage_tiles = splits == 1 ? 2 * 32 * BM * 4 : 0;
    const int avail = 227 * 1024 - 4096 - stage_tiles;
    auto fits = [&](int taps, int min_stages) {
        const int sb = (taps == 9 ? 23 * 1024 : 20 * 1024) + taps * (bn / 2) * BK_BYTES;
        int stages = std::min(avail / sb, MAX_STAGES);
        if (splits > 1) {
            while ((size_t)stages * sb < (size_t)bn * BM * 4) ++stages;
            if ((size_t)stages * sb + 2048 > 227 * 1024 - 2048 || stages > MAX_STAGES) return false;
        }
        return stages >= min_stages;
    };
    if (fits(9, 2)) return 9;
    if (fits(3, 3)) return 3;
    return 0;
}

int b200_launch_conv_tc2(cudaStream_t s, const b200_device_info& dev, const b200_conv_args& c, int bn, int splits, int halo_taps) {
    if (!b200_conv_tc_supported(c.N, c.H, c.W, c.C, c.OC, c.KH, c.KW, 1, 1, c.pad, c.pad, c.dil, c.dil)) return -1;
    if (((uintptr_t)c.x_nhwc & 15) || ((uintptr_t)c.w_packed & 15)) return -1;
    if (bn < 16 || bn > 256 || (bn & 15) || splits < 1 || splits > 4) return -1;
    const int64_t M = c.H * c.W, K = (int64_t)c.KH * c.KW * c.C;
    // a pair tile is 256 consecutive output pixels of one image: the image must split into whole 128-pixel boxes
    if (M % 128 != 0) return -1;
    if (halo_taps) {
        // 3x3, stride 1, pad 1; 16 x 8 pixel patches tile the image
        if ((halo_taps != 3 && halo_taps != 9) || c.KH != 3 || c.KW != 3 || c.pad != 1 || c.dil != 1 || c.W % 8 || c.H % 16) return -1;
    }
    const int cblocks = (int)(c.C / 64);
    const int nkb = halo_taps ? (halo_taps == 9 ? cblocks : 3 * cblocks) : (int)(K / 64);      // ring stages per tile
    if (splits > nkb) return -1;
    G2Params kp;
    memset(&kp, 0, sizeof(kp));
    unsigned ctas = 0;
    size_t smem = 0;
    if (!fill_geometry(kp, dev, M, c.OC, c.N, nkb, bn, splits, &ctas, &smem, halo_taps)) return -1;
    const uint32_t BW = halo_taps ? 10u : (uint32_t)(c.W < 128 ? c.W : 128), BH = halo_taps ? (halo_taps == 9 ? 18u : 16u) : 128 / BW;
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
    if (!encode_rows(&tb, c.w_packed, GGML_TYPE_F16, K, c.OC, K, 1, 0, (uint32_t)(bn / 2))) return -1;
    kp.D = c.D; kp.ldd = M; kp.d_batch_stride = c.OC * M;
    kp.M = M; kp.N = c.OC;
    kp.ne12 = (int)c.N; kp.r2 = 1;
    if (c.w_prefetch) { kp.wpf = (const char*)c.w_packed; kp.wpf_ld = K * 2; kp.wpf_rows = c.OC; kp.wpf_kbytes = K * 2; kp.wpf_is_a = 0; }
    kp.bias = c.bias; kp.bias_mode = c.bias ? 2 : 0;
    kp.residual = c.residual; kp.ldr = M; kp.r_batch_stride = c.OC * M;
    kp.D2 = c.D2; kp.d2_seq = c.d2_seq; kp.d2_slot = c.d2_slot_floats;
    kp.vec_epi = vec_epilogue_ok(kp, splits);
    if (halo_taps && !kp.vec_epi && splits == 1) return -1;        // the per-row epilogues do not know the 16 x 8 patch layout
    if (halo_taps) { kp.wpf = nullptr; }                            // (the up-front L2 request counts k-blocks of 128 bytes, not stages)
    kp.conv = 1; kp.conv_W = (int)c.W; kp.conv_KW = c.KW; kp.conv_cblocks = (int)(c.C / 64); kp.conv_pad = c.pad; kp.conv_dil = c.dil;
    CUtensorMap td;
    encode_output(&td, kp, c.N);
    if (halo_taps) kp.tma_store = 0;
    cudaError_t e = launch2<0>(s, ctas, (unsigned)splits, smem, ta, tb, td, kp);
    if (e != cudaSuccess) {
        fprintf(stderr, "[ggml-b200] CTA-pair conv launch failed: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
        return -1;
    }
    return 1;
}
