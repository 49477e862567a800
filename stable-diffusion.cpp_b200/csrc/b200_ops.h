// b200_ops.h -- launchers of the hand-written sm_100a kernels (kernels/*.cu).
// Plain C++ signatures: raw device pointers + byte strides in b200_td, a stream, scalars.
// Every launcher enqueues on `stream` and returns the number of kernels it launched.
#pragma once

#include "b200_common.h"

// ---- elementwise.cu ------------------------------------------------------------------------
enum b200_binop { B200_ADD = 0, B200_SUB = 1, B200_MUL = 2, B200_DIV = 3 };
int b200_launch_binary(cudaStream_t s, int op, const b200_td& a, const b200_td& b, const b200_td& dst);
// dst = x * gelu_tanh(gate) (f32, row-vectorisable views); optional contiguous f16 copy of dst for the contraction that follows
int b200_launch_geglu(cudaStream_t s, const b200_td& x, const b200_td& gate, const b200_td& dst, void* dst16);
// unary ops use ggml_unary_op numbering; p0/p1 are op parameters (unused by most)
int b200_launch_unary(cudaStream_t s, int unary_op, const b200_td& src, const b200_td& dst);
// extra scalar ops that are separate GGML_OPs
enum b200_scalar_op { B200_SCALE = 0, B200_CLAMP = 1, B200_SQR = 2, B200_SQRT = 3, B200_LEAKY_RELU = 4, B200_SIN = 5, B200_COS = 6, B200_LOG = 7 };
int b200_launch_scalar_op(cudaStream_t s, int op, const b200_td& src, const b200_td& dst, float p0, float p1);
// strided, type-converting copy (CPY / CONT / DUP); src and dst have equal element counts
int b200_launch_copy(cudaStream_t s, const b200_td& src, const b200_td& dst);
int b200_launch_concat(cudaStream_t s, const b200_td& a, const b200_td& b, const b200_td& dst, int dim);
int b200_launch_repeat(cudaStream_t s, const b200_td& src, const b200_td& dst);
int b200_launch_pad(cudaStream_t s, const b200_td& src, const b200_td& dst, const int32_t* pads /*lp0,rp0,lp1,rp1,lp2,rp2,lp3,rp3*/, bool circular);
int b200_launch_upscale(cudaStream_t s, const b200_td& src, const b200_td& dst, int mode_flags);
int b200_launch_timestep_embedding(cudaStream_t s, const b200_td& src, const b200_td& dst, int dim, int max_period);
int b200_launch_get_rows(cudaStream_t s, const b200_td& src, const b200_td& idx, const b200_td& dst);
int b200_launch_arange(cudaStream_t s, const b200_td& dst, float start, float step);
int b200_launch_fill(cudaStream_t s, const b200_td& dst, float value);
int b200_launch_glu(cudaStream_t s, int glu_op, const b200_td& a, const b200_td* b, const b200_td& dst, bool swapped);
int b200_launch_sum_rows(cudaStream_t s, const b200_td& src, const b200_td& dst, bool mean);
// f32 -> f16/bf16 pack of a strided 2-D/4-D operand into a dense K-major matrix with K padded to kpad (zero filled)
int b200_launch_pack_rows(cudaStream_t s, const b200_td& src, void* dst, int dst_type, int64_t kpad);

// activation of a Q8_0-weight contraction: quantise each 32-value block like the oracle (quantize_row_q8_0) and store d * q as f16
int b200_launch_pack_rows_q8_roundtrip(cudaStream_t s, const b200_td& a, void* dst_f16, int64_t kpad);
// 3xTF32: f32 [K, rows, b2, b3] (any strides) -> hi / lo dense [rows][kpad] f32 (x = hi + lo, hi exactly representable in TF32)
int b200_launch_split_tf32(cudaStream_t s, const b200_td& a, float* hi, float* lo, int64_t kpad);

// 16-bit tiled transpose: src [d, L, b2, b3] (unit stride along d) -> dst dense [b3][b2][d rows][Lpad] (row = d index, L contiguous)
int b200_launch_transpose_f16(cudaStream_t s, const b200_td& src, void* dst, int64_t Lpad);

// ---- norm.cu ---------------------------------------------------------------------------------
// optional fused affine (w, b: per-channel f32, may be null) and activation (0 none, 1 SiLU)
int b200_launch_group_norm(cudaStream_t s, const b200_td& src, const b200_td& dst, int n_groups, float eps, const float* w = nullptr,
                           const float* b = nullptr, int act = 0);
enum b200_norm_kind { B200_NORM_LAYER = 0, B200_NORM_RMS = 1, B200_NORM_L2 = 2 };
// optional fused affine (w, b: per-column f32 of length ne0, may be null)
// out16 (optional): additionally write the result rounded to out16_type (F16 / BF16) as a dense [rows][ne0] matrix
int b200_launch_norm(cudaStream_t s, int kind, const b200_td& src, const b200_td& dst, float eps, const float* w = nullptr, const float* b = nullptr,
                     void* out16 = nullptr, int out16_type = -1, int modulate = 0);
// modulate = 1: y = (n + n * w) + b with every step rounded (adaLN `x * (1 + scale) + shift` as the reference graph spells it: MUL, ADD, ADD)
int b200_launch_soft_max(cudaStream_t s, const b200_td& src, const b200_td* mask, const b200_td& dst, float scale, float max_bias);

// ---- im2col.cu -------------------------------------------------------------------------------
// p = op_params {s0,s1,s2,p0,p1,p2,d0,d1,d2}; src [IW,IH,ID,N*IC] f32 -> dst [IC*KD*KH*KW, OW, OH, N*OD] f16|f32
int b200_launch_im2col_3d(cudaStream_t s, const b200_td& src, const b200_td& dst, int64_t KW, int64_t KH, int64_t KD, int64_t IC, const int32_t* p);
int b200_launch_im2col(cudaStream_t s, const b200_td& src /*image*/, const b200_td& dst, int64_t KW, int64_t KH, int s0, int s1, int p0,
                       int p1, int d0, int d1, bool is_2d);

// ---- gemm_ref.cu: CUDA-core GEMM (bring-up / odd shapes) ------------------------------------------
// dst[n][m] (f32, m fastest, ldd floats between n) = sum_k A[m][k] * B[n][k];  A,B rows K-contiguous, type f32/f16/bf16
int b200_launch_gemm_ref(cudaStream_t s, const void* A, int a_type, int64_t lda_bytes, const void* B, int b_type, int64_t ldb_bytes,
                         float* D, int64_t ldd, int64_t M, int64_t N, int64_t K);

// ---- gemm_tc.cu: tcgen05 / TMEM / TMA GEMM -------------------------------------------------------
struct b200_gemm_args {
    const void* A;          // [M rows][K] K-major, element type `type` (GGML_TYPE_F16 / BF16 / F32 -> kind::f16 / kind::f16(bf16) / kind::tf32)
    const void* B;          // [N rows][K] K-major, same element type
    int         type;
    int64_t     M, N, K;
    int64_t     lda, ldb;   // row strides in ELEMENTS (multiple of 16 bytes)
    int64_t     batch;      // number of independent problems (grid.z); strides below in elements
    int64_t     a_batch_stride, b_batch_stride, d_batch_stride;
    int64_t     a_bcast;    // batch index of A = z / a_bcast (ggml broadcast of src0 over src1 batches); >= 1
    float*      D;          // [N][M] f32, M fastest (dst->data), ldd floats between consecutive n
    int64_t     ldd;
    const float* bias;      // optional, nullptr if none
    int         bias_mode;  // 0 none, 1 per-m (row of A: linear bias), 2 per-n (row of B: conv bias / channel)
    const float* residual;  // optional [N][M] like D (added after bias)
    int64_t     ldr;
    const float* gate;      // optional per-m f32 vector: D = residual + gate[m] * act(acc + bias) -- the gated residual of the DiT blocks
                            // (x + gate * Linear(y): flux.hpp DoubleStreamBlock / SingleStreamBlock); product and sum rounded separately
    int         act;        // 0 none, 1 SiLU, 2 GELU(tanh)
    int         early;      // bit 0: A, bit 1: B is a constant (weight) operand no kernel of this graph writes -> may be fetched before the PDL wait
    int         wprefetch;  // bit 0: A, bit 1: B is a constant weight operand (batch 1): every CTA requests the whole slab of weight rows it will
                            // stage (its rows x its K range) from L2 with cp.async.bulk.prefetch.L2 BEFORE the PDL wait, so the HBM latency of a
                            // weight stream that is read exactly once per forward is paid up front instead of per ring slot (option wprefetch)
    void*       trace;      // optional device buffer of 8 uint64: phase timestamps of CTA (0,0,0) (tools/gemm_bench)
    // optional 16-bit copy of the result (same [N][M] element layout, ldd / d_batch_stride in elements): the K-major operand of the
    // contraction that consumes it, rounded exactly like its operand pack would.  Honoured by the CTA-pair kernel's staged epilogue only:
    // *d16_done is set to 1 when it was written.  skip_f32: the f32 tensor has no other reader and need not be stored at all
    // (f32 output bytes are what bounds the MLP-up projections: ~2.7 TB/s of write bandwidth on this part).
    void*       D16;
    int         d16_type;   // GGML_TYPE_F16 | GGML_TYPE_BF16
    int         skip_f32;
    int*        d16_done;
    int64_t     geglu;      // > 0: GEGLU mode of the CTA-pair kernel (block.hpp:182-210).  A = [2 * geglu rows][K]: x-half features then gate-half; the
                            // ONLY output is D16 = x * gelu_tanh(gate) as [N tokens][geglu] 16-bit rows (batch stride N * geglu); M == 2 * geglu,
                            // geglu % 64 == 0, bias per row or none, skip_f32 set.  The launcher returns -1 when it cannot run it (nothing launched)
    int         d16_strict; // return -1 WITHOUT launching when the 16-bit copy cannot be written (the caller depends on it: a projection run ahead
                            // of its turn on a side stream must not touch its f32 tensor, whose memory still belongs to somebody else)
};
// returns kernels launched, or -1 if the shape/alignment is not supported by the TMA path (caller falls back)
int b200_launch_gemm_tc(cudaStream_t s, const b200_device_info& dev, const b200_gemm_args& g, void* workspace, size_t workspace_bytes);
size_t b200_gemm_tc_workspace_bytes(const b200_device_info& dev, const b200_gemm_args& g);

// gemm_tc2.cu: CTA-pair (tcgen05 cta_group::2, M = 256) persistent GEMM / implicit conv with two TMEM accumulators; F16 / BF16 only.
// bn: tile N of the pair (multiple of 16, <= 256), splits: split-K factor inside the cluster (1..4).  1 when launched, -1 when the
// problem is outside the envelope.  The conv front end is declared after b200_conv_args below.
int b200_launch_gemm_tc2(cudaStream_t s, const b200_device_info& dev, const b200_gemm_args& g, int bn, int splits);
// a_bytes: bytes of A one CTA stages per 64-wide k-block (16 KB; 20 KB / 3 or 22.5 KB / 9 for the halo-reuse convolution)
double b200_gemm_tc2_model(const b200_device_info& dev, int64_t M, int64_t N, int64_t batch, int nkb, int bn, int splits, double a_bytes = 16384.0);

// Q8_0 blocks (34 bytes: f16 scale + 32 int8, ggml-common.h:251-255) -> f16 rows [rows][K] (K % 32 == 0): the derived weight layout the
// tensor-core GEMM reads; value = round_f16(float(d) * q), the reference's own dequantisation (ggml-quants.c dequantize_row_q8_0)
int b200_launch_dequant_q8_0(cudaStream_t s, const void* blocks, void* out_f16, int64_t n_blocks);

// ---- rope.cu: Rope::apply_rope (interleaved) as one pass, optional RMSNorm*scale prologue, f32 or f16 rows out ------------------
int b200_launch_rope(cudaStream_t s, const b200_td& x, const float* pe, void* out, int out_type, const float* rms_w, float eps);

// ---- gemv.cu: MUL_MAT with N <= 4 activation rows, F16/BF16 weights read in place; pre_act 1 = SiLU on the activation -----
bool b200_gemv_supported(int wtype, int64_t M, int64_t N, int64_t K, const void* W, int64_t lda, const void* X);
int b200_launch_gemv(cudaStream_t s, int wtype, const void* W, int64_t lda, const float* X, int64_t ldx, float* D, int64_t ldd, int64_t M, int64_t N,
                     int64_t K, const float* bias, const float* residual, int64_t ldr, int pre_act);

// ---- implicit-GEMM convolution (gemm_tc.cu conv mode + conv_prep.cu operand producers) ------------------------
struct b200_conv_args {
    const void* x_nhwc;     // f16 [N][H][W][C]   (produced by b200_launch_to_nhwc_f16)
    const void* w_packed;   // f16 [OC][KH][KW][C] (produced by b200_launch_pack_conv_weight)
    int64_t N, H, W, C, OC;
    int KH, KW, pad, dil;
    float* D;               // f32 [N][OC][H][W] == ggml [W,H,OC,N]
    const float* bias;      // per OC or null
    const float* residual;  // same layout as D or null
    int w_const;            // w_packed was not produced by a kernel of this graph execution (may be fetched before the PDL wait)
    int w_prefetch;         // request each CTA's filter slab from L2 up front (see b200_gemm_args::wprefetch); implies w_const
    // optional second destination over NVLink (kernels/peer.cu): every output element is also stored at
    // D2 + ((*d2_seq + 1) & 1) * d2_slot_floats + (its offset in D).  Honoured by the CTA-pair kernel only (launcher returns 2).
    float* D2;
    const unsigned* d2_seq;
    int64_t d2_slot_floats;
};
bool b200_conv_tc_supported(int64_t N, int64_t H, int64_t W, int64_t C, int64_t OC, int KH, int KW, int s0, int s1, int p0, int p1, int d0, int d1);
size_t b200_conv_tc_workspace_bytes(const b200_device_info& dev, const b200_conv_args& c);
int b200_launch_conv_tc(cudaStream_t s, const b200_device_info& dev, const b200_conv_args& c, void* workspace, size_t workspace_bytes);
// halo_taps 3 | 9: 3x3 convolution with the image box staged once per 3 | 9 taps (16 x 8 pixel patches; see gemm_tc2.cu), 0: per-tap boxes
int b200_launch_conv_tc2(cudaStream_t s, const b200_device_info& dev, const b200_conv_args& c, int bn, int splits, int halo_taps = 0);
int b200_conv_tc2_halo_taps(int bn, int splits);
// stats: float2 {mean, rstd} per (image, group)
// partial / counters (optional): scratch of b200_gn_stats_partial_bytes() and B200_GN_COUNTERS zero-initialised unsigneds owned by the
// backend instance: large groups are then split over several CTAs (one read of x, deterministic merge by the last CTA)
#define B200_GN_COUNTERS 4096
// addv (optional): per-(image, channel) f32 vector [N][C] added to x on the fly (the ResBlock's `h + emb` broadcast ADD, block.hpp:150-160)
int b200_launch_gn_stats(cudaStream_t s, const float* x, float* stats, int64_t N, int64_t C, int64_t inner, int n_groups, float eps, void* partial = nullptr,
                         unsigned* counters = nullptr, const float* addv = nullptr);
size_t b200_gn_stats_partial_bytes(int64_t N, int64_t C, int64_t inner, int n_groups);
// NCHW f32 -> NHWC f16 with optional GroupNorm (stats + per-channel w, b), SiLU (act = 1) and nearest upsampling (up = 1 | 2)
int b200_launch_to_nhwc_f16(cudaStream_t s, const float* x, void* out, int64_t N, int64_t C, int64_t H, int64_t W, int up, const float* stats,
                            int n_groups, const float* gw, const float* gb, int act, const float* addv = nullptr);
int b200_launch_pack_conv_weight(cudaStream_t s, const void* w, void* out, int KW, int KH, int64_t IC, int64_t OC);

// ---- peer.cu: CFG-split exchange over NVLink peer memory -------------------------------------------------------------------------
int b200_launch_peer_push(cudaStream_t s, const void* src, void* peer_base, const unsigned* seq, size_t bytes, size_t slot_bytes);
int b200_launch_peer_signal_wait(cudaStream_t s, unsigned* my_seq, unsigned* peer_flag, const unsigned* my_flag, unsigned* my_err, double timeout_s);

// ---- attention.cu --------------------------------------------------------------------------------
// ggml FLASH_ATTN_EXT: q f32 [d, Lq, H, N], k f16 [d, Lk, Hkv, N], v f16 [dv, Lk, Hkv, N], mask f16 [Lk, >=Lq, ...] or null,
// dst f32 [dv, H, Lq, N]
// vt = packed V^T f16 [Lk_pad, dv, Hkv, N]; returns -1 when the shape is outside the fused kernel's envelope.  Only the STRIDES of q / k /
// v / dst are read besides q's and k's extents, so a caller may describe projections that were never permuted into ggml's layout (heads
// interleaved inside a token row, the batch as its own dimension).  dst16: f16 copy of the result with dst's element layout;
// skip_f32: write only that copy (its one reader is the output projection)
int b200_launch_flash_attn_fused(cudaStream_t s, const b200_td& q, const b200_td& k, const void* vt, int64_t Lk_pad, const b200_td& v,
                                 const b200_td* mask, const b200_td& dst, float scale, void* dst16 = nullptr, int skip_f32 = 0);
