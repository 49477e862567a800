// b200_graph.cpp -- graph_compute for the B200 backend: walks the ggml graph handed over by the reference's
// host code (ggml_backend_graph_compute, ggml/src/ggml-backend.cpp:444-452) and turns nodes into launches of the
// hand-written kernels in kernels/*.cu.  No node ever leaves the GPU: an op we cannot run is reported by
// supports_op() at graph-build time, never silently computed elsewhere.
//
// Node walk contract (same as the reference backends, ggml-cpu.c:3255-3287): nodes are topologically ordered;
// RESHAPE / VIEW / PERMUTE / TRANSPOSE / NONE are metadata only; nodes without GGML_TENSOR_FLAG_COMPUTE are skipped.

#include "b200_graph.h"
#include "b200_ops.h"

#include "ggml-backend-impl.h"
#include "ggml-impl.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
static int env_flag(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}

b200_context* b200_context_create(const b200_device_info& info) {
    auto* ctx = new b200_context;
    ctx->device = info.id;
    ctx->info = info;
    ctx->name = info.name;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return nullptr;
    }
    cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming);
    cudaEventCreate(&ctx->ev_start);
    cudaEventCreate(&ctx->ev_stop);
    ctx->opt_fusion = env_flag("GGML_B200_FUSION", 1) != 0;
    ctx->opt_tc_gemm = env_flag("GGML_B200_TC_GEMM", 1) != 0;
    ctx->opt_timing = env_flag("GGML_B200_TIMING", 1) != 0;
    ctx->opt_cuda_graphs = env_flag("GGML_B200_CUDA_GRAPHS", 0) != 0;
    return ctx;
}

b200_context::~b200_context() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto& c : ws.chunks) cudaFree(c.base);
    for (auto& p : kt_pending) { cudaEventDestroy(p.start); cudaEventDestroy(p.stop); }
    for (auto e : kt_free) cudaEventDestroy(e);
    if (copy_event) cudaEventDestroy(copy_event);
    if (ev_start) cudaEventDestroy(ev_start);
    if (ev_stop) cudaEventDestroy(ev_stop);
    if (stream) cudaStreamDestroy(stream);
}

int b200_context_set_option(b200_context* ctx, const char* key, int value) {
    if (!strcmp(key, "fusion")) ctx->opt_fusion = value != 0;
    else if (!strcmp(key, "tc_gemm")) ctx->opt_tc_gemm = value != 0;
    else if (!strcmp(key, "timing")) ctx->opt_timing = value != 0;
    else if (!strcmp(key, "cuda_graphs")) ctx->opt_cuda_graphs = value != 0;
    else if (!strcmp(key, "kernel_timing")) ctx->opt_kernel_timing = value != 0;
    else return -1;
    return 0;
}

static void kt_flush(b200_context* ctx);

void b200_context_finalize_timing(b200_context* ctx) {
    kt_flush(ctx);
    if (!ctx->timing_pending) return;
    cudaSetDevice(ctx->device);
    if (cudaEventSynchronize(ctx->ev_stop) == cudaSuccess) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop);
        ctx->stats.last_graph_ms = ms;
        ctx->stats.total_graph_ms += ms;
    } else {
        cudaGetLastError();
    }
    ctx->timing_pending = false;
}

void b200_invalidate_address_range(int, const void*, size_t) {
    // derived-layout caches (packed weights) are introduced together with the fused conv path; nothing cached yet
}

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
static void ws_begin_graph(b200_context* ctx) {
    auto& ws = ctx->ws;
    if (ws.chunks.size() > 1) {
        // consolidate: previous graph overflowed into extra chunks; drain the stream and get one block
        size_t total = 0;
        for (auto& c : ws.chunks) total += c.size;
        B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (auto& c : ws.chunks) cudaFree(c.base);
        ws.chunks.clear();
        char* p = nullptr;
        if (cudaMalloc(&p, total) == cudaSuccess) ws.chunks.push_back({p, total, 0});
        else cudaGetLastError();
    }
    for (auto& c : ws.chunks) c.used = 0;
    ws.high_water = 0;
}

static void* ws_alloc(b200_context* ctx, size_t bytes) {
    auto& ws = ctx->ws;
    bytes = (bytes + 1023) & ~(size_t)1023;
    ws.high_water += bytes;
    if (!ws.chunks.empty()) {
        auto& c = ws.chunks.back();
        if (c.used + bytes <= c.size) {
            void* p = c.base + c.used;
            c.used += bytes;
            return p;
        }
    }
    size_t sz = std::max(bytes, (size_t)64 << 20);
    if (!ws.chunks.empty()) sz = std::max(sz, ws.chunks.back().size);
    char* p = nullptr;
    if (cudaMalloc(&p, sz) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    ws.chunks.push_back({p, sz, bytes});
    return p;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static inline bool is_f32(const ggml_tensor* t) { return t->type == GGML_TYPE_F32; }
static inline bool is_fp(const ggml_tensor* t) { return t->type == GGML_TYPE_F32 || t->type == GGML_TYPE_F16 || t->type == GGML_TYPE_BF16; }
static inline int64_t fp_size(int type) { return type == GGML_TYPE_F32 ? 4 : 2; }

static inline bool rows_unit_stride(const ggml_tensor* t) { return t->nb[0] == ggml_type_size(t->type); }

// can the TMA describe this K-major operand in place?  (16-byte aligned base and strides, unit stride along K)
static bool tma_compatible(const ggml_tensor* t) {
    if (!rows_unit_stride(t)) return false;
    if ((uintptr_t)t->data % 16) return false;
    if (t->nb[1] % 16) return false;
    if (t->ne[2] > 1 && t->nb[2] % 16) return false;
    if (t->ne[3] > 1 && t->nb[3] % 16) return false;
    return true;
}

struct operand {
    const void* ptr;
    int type;
    int64_t ld;            // row stride, elements
    int64_t batch_stride;  // dim-2 stride, elements
    int64_t b3_stride;     // dim-3 stride, elements
};

// Bring `t` ([K, rows, b2, b3]) into a form the tcgen05 GEMM can read as type `want`: in place when possible,
// otherwise packed (converted, K padded to 16 bytes) into workspace.  Returns false on allocation failure.
static bool prepare_operand(b200_context* ctx, const ggml_tensor* t, int want, operand* out, int* launches) {
    if ((int)t->type == want && tma_compatible(t)) {
        const int64_t es = fp_size(want);
        *out = operand{t->data, want, (int64_t)t->nb[1] / es, (int64_t)t->nb[2] / es, (int64_t)t->nb[3] / es};
        return true;
    }
    const int64_t es = fp_size(want);
    const int64_t kal = 16 / es;
    const int64_t kpad = (t->ne[0] + kal - 1) / kal * kal;
    const int64_t rows = t->ne[1] * t->ne[2] * t->ne[3];
    void* buf = ws_alloc(ctx, (size_t)(rows * kpad * es));
    if (!buf) return false;
    int n = b200_launch_pack_rows(ctx->stream, b200_make_td(t), buf, want, kpad);
    if (n < 0) return false;
    *launches += n;
    *out = operand{buf, want, kpad, kpad * t->ne[1], kpad * t->ne[1] * t->ne[2]};
    return true;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 GEMM launch wrapper: workspace for split-K, counters, optional per-launch CUDA-event timing
// (option "kernel_timing": the roofline numerator/denominator bench.py reports for the dominant kernel)
// ------------------------------------------------------------------------------------------------
static void kt_flush(b200_context* ctx) {
    if (ctx->kt_pending.empty()) return;
    cudaSetDevice(ctx->device);
    for (auto& p : ctx->kt_pending) {
        float ms = 0;
        if (cudaEventSynchronize(p.stop) == cudaSuccess && cudaEventElapsedTime(&ms, p.start, p.stop) == cudaSuccess) {
            ctx->kt_us += (double)ms * 1e3;
            ctx->kt_flops += p.flops;
        } else {
            cudaGetLastError();
        }
        ctx->kt_free.push_back(p.start);
        ctx->kt_free.push_back(p.stop);
    }
    ctx->kt_pending.clear();
    ctx->stats.reserved[0] = (uint64_t)ctx->kt_flops;
    ctx->stats.reserved[1] = (uint64_t)ctx->kt_us;
}

static cudaEvent_t kt_event(b200_context* ctx) {
    if (!ctx->kt_free.empty()) { cudaEvent_t e = ctx->kt_free.back(); ctx->kt_free.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

static int launch_tc(b200_context* ctx, const b200_gemm_args& g) {
    if (!ctx->opt_tc_gemm) return -1;
    size_t wsb = b200_gemm_tc_workspace_bytes(ctx->info, g);
    void* w = wsb ? ws_alloc(ctx, wsb) : nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->opt_kernel_timing) {
        e0 = kt_event(ctx);
        e1 = kt_event(ctx);
        cudaEventRecord(e0, ctx->stream);
    }
    int n = b200_launch_gemm_tc(ctx->stream, ctx->info, g, w, w ? wsb : 0);
    if (ctx->opt_kernel_timing) {
        if (n > 0) {
            cudaEventRecord(e1, ctx->stream);
            ctx->kt_pending.push_back({e0, e1, 2.0 * (double)g.M * (double)g.N * (double)g.K * (double)g.batch});
        } else {
            ctx->kt_free.push_back(e0);
            ctx->kt_free.push_back(e1);
        }
    }
    if (n > 0) ctx->stats.tc_gemm_launches += n;
    return n;
}

// ------------------------------------------------------------------------------------------------
// MUL_MAT
// ------------------------------------------------------------------------------------------------
static int compute_type_for(const ggml_tensor* src0) {
    // the CPU oracle converts src1 to src0's vec_dot type (ggml-cpu.c:1430-1513): f16 weights -> f16 x f16 with f32
    // accumulation, bf16 -> bf16 x bf16, f32 -> f32 (here: tf32 tensor-core inputs, f32 accumulation)
    if (src0->type == GGML_TYPE_F16) return GGML_TYPE_F16;
    if (src0->type == GGML_TYPE_BF16) return GGML_TYPE_BF16;
    return GGML_TYPE_F32;
}

static int op_mul_mat(b200_context* ctx, ggml_tensor* dst) {
    const ggml_tensor* src0 = dst->src[0];
    const ggml_tensor* src1 = dst->src[1];
    int launches = 0;
    const int ct = compute_type_for(src0);
    const int64_t K = src0->ne[0], M = src0->ne[1], N = src1->ne[1];
    const int64_t ne02 = src0->ne[2], ne03 = src0->ne[3], ne12 = src1->ne[2], ne13 = src1->ne[3];
    const int64_t r2 = ne12 / ne02, r3 = ne13 / ne03;
    if (M == 0 || N == 0 || ne12 * ne13 == 0) return 0;
    if (K == 0) {
        cudaMemsetAsync(dst->data, 0, ggml_nbytes(dst), ctx->stream);
        return 1;
    }

    operand a, b;
    if (!prepare_operand(ctx, src0, ct, &a, &launches)) return -1;
    if (!prepare_operand(ctx, src1, ct, &b, &launches)) return -1;

    for (int64_t i3 = 0; i3 < ne13; ++i3) {
        const int64_t es = fp_size(ct);
        b200_gemm_args g;
        memset(&g, 0, sizeof(g));
        g.A = (const char*)a.ptr + (i3 / r3) * a.b3_stride * es;
        g.B = (const char*)b.ptr + i3 * b.b3_stride * es;
        g.type = ct;
        g.M = M; g.N = N; g.K = K;
        g.lda = a.ld; g.ldb = b.ld;
        g.batch = ne12;
        g.a_batch_stride = a.batch_stride;
        g.b_batch_stride = b.batch_stride;
        g.a_bcast = r2;
        g.D = (float*)((char*)dst->data + i3 * dst->nb[3]);
        g.ldd = dst->nb[1] / 4;
        g.d_batch_stride = dst->nb[2] / 4;
        int n = launch_tc(ctx, g);
        if (n < 0) {
            // CUDA-core reference kernel (debug option, or shapes the TMA cannot describe)
            n = 0;
            for (int64_t i2 = 0; i2 < ne12; ++i2) {
                int r = b200_launch_gemm_ref(ctx->stream, (const char*)g.A + (i2 / r2) * a.batch_stride * es, ct, a.ld * es,
                                             (const char*)g.B + i2 * b.batch_stride * es, ct, b.ld * es, g.D + i2 * g.d_batch_stride, g.ldd, M, N, K);
                if (r < 0) return -1;
                n += r;
            }
        }
        launches += n;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// FLASH_ATTN_EXT (v1: tensor-core GEMMs + row softmax through workspace; the fused single-kernel
// version replaces this for the head sizes it covers)
// ------------------------------------------------------------------------------------------------
static int op_flash_attn(b200_context* ctx, ggml_tensor* dst) {
    const ggml_tensor* q = dst->src[0];
    const ggml_tensor* k = dst->src[1];
    const ggml_tensor* v = dst->src[2];
    const ggml_tensor* mask = dst->src[3];
    float scale, max_bias;
    memcpy(&scale, (const float*)dst->op_params + 0, sizeof(float));
    memcpy(&max_bias, (const float*)dst->op_params + 1, sizeof(float));
    const int64_t d = q->ne[0], Lq = q->ne[1], H = q->ne[2], NB = q->ne[3];
    const int64_t Lk = k->ne[1], Hkv = k->ne[2], dv = v->ne[0];
    const int64_t rk = H / Hkv;
    int launches = 0;
    if (Lq * H * NB == 0) return 0;
    const int ct = k->type == GGML_TYPE_BF16 ? GGML_TYPE_BF16 : GGML_TYPE_F16;
    const int64_t es = 2;
    const int64_t Lk_pad = (Lk + 7) / 8 * 8;

    operand qa, ka;
    if (!prepare_operand(ctx, q, ct, &qa, &launches)) return -1;
    if (!prepare_operand(ctx, k, ct, &ka, &launches)) return -1;
    // V^T: [Lk, dv, Hkv, NB] view of v, packed to ct with Lk padded
    ggml_tensor vt = *v;
    vt.ne[0] = v->ne[1]; vt.nb[0] = v->nb[1];
    vt.ne[1] = v->ne[0]; vt.nb[1] = v->nb[0];
    void* vbuf = ws_alloc(ctx, (size_t)(Lk_pad * dv * Hkv * NB * es));
    float* sbuf = (float*)ws_alloc(ctx, (size_t)(Lk * Lq * H * sizeof(float)));
    void* pbuf = ws_alloc(ctx, (size_t)(Lk_pad * Lq * H * es));
    if (!vbuf || !sbuf || !pbuf) return -1;
    int n = b200_launch_pack_rows(ctx->stream, b200_make_td(&vt), vbuf, ct, Lk_pad);
    if (n < 0) return -1;
    launches += n;

    for (int64_t i3 = 0; i3 < NB; ++i3) {
        // S[Lk, Lq, H] = K . Q^T
        b200_gemm_args g;
        memset(&g, 0, sizeof(g));
        g.A = (const char*)ka.ptr + i3 * ka.b3_stride * es;
        g.B = (const char*)qa.ptr + i3 * qa.b3_stride * es;
        g.type = ct;
        g.M = Lk; g.N = Lq; g.K = d;
        g.lda = ka.ld; g.ldb = qa.ld;
        g.batch = H;
        g.a_batch_stride = ka.batch_stride;
        g.b_batch_stride = qa.batch_stride;
        g.a_bcast = rk;
        g.D = sbuf; g.ldd = Lk; g.d_batch_stride = Lk * Lq;
        n = launch_tc(ctx, g);
        if (n < 0) {
            n = 0;
            for (int64_t h = 0; h < H; ++h)
                n += b200_launch_gemm_ref(ctx->stream, (const char*)g.A + (h / rk) * ka.batch_stride * es, ct, ka.ld * es,
                                          (const char*)g.B + h * qa.batch_stride * es, ct, qa.ld * es, sbuf + h * Lk * Lq, Lk, Lk, Lq, d);
        }
        launches += n;

        // P = softmax(S * scale + mask) (f32, in place)
        b200_td std_;
        std_.data = sbuf; std_.type = GGML_TYPE_F32;
        std_.ne[0] = Lk; std_.ne[1] = Lq; std_.ne[2] = H; std_.ne[3] = 1;
        std_.nb[0] = 4; std_.nb[1] = Lk * 4; std_.nb[2] = Lk * Lq * 4; std_.nb[3] = Lk * Lq * H * 4;
        b200_td mtd;
        if (mask) {
            mtd = b200_make_td(mask);
            mtd.data = (char*)mask->data + (i3 % mask->ne[3]) * mask->nb[3];
            mtd.ne[3] = 1;
        }
        n = b200_launch_soft_max(ctx->stream, std_, mask ? &mtd : nullptr, std_, scale, max_bias);
        if (n < 0) return -1;
        launches += n;
        // P -> f16 [H][Lq][Lk_pad]
        n = b200_launch_pack_rows(ctx->stream, std_, pbuf, ct, Lk_pad);
        if (n < 0) return -1;
        launches += n;

        // O[dv, H, Lq] = V^T . P^T  (per head; written straight into ggml's [dv, H, Lq, N] layout)
        memset(&g, 0, sizeof(g));
        g.A = (const char*)vbuf + i3 * (Lk_pad * dv * Hkv) * es;
        g.B = pbuf;
        g.type = ct;
        g.M = dv; g.N = Lq; g.K = Lk;
        g.lda = Lk_pad; g.ldb = Lk_pad;
        g.batch = H;
        g.a_batch_stride = Lk_pad * dv;
        g.b_batch_stride = Lk_pad * Lq;
        g.a_bcast = rk;
        g.D = (float*)((char*)dst->data + i3 * dst->nb[3]);
        g.ldd = dst->nb[2] / 4;       // between consecutive queries
        g.d_batch_stride = dst->nb[1] / 4;   // between heads
        n = launch_tc(ctx, g);
        if (n < 0) {
            n = 0;
            for (int64_t h = 0; h < H; ++h)
                n += b200_launch_gemm_ref(ctx->stream, (const char*)g.A + (h / rk) * g.a_batch_stride * es, ct, Lk_pad * es,
                                          (const char*)pbuf + h * g.b_batch_stride * es, ct, Lk_pad * es, g.D + h * g.d_batch_stride, g.ldd, dv, Lq, Lk);
        }
        launches += n;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// supports_op
// ------------------------------------------------------------------------------------------------
bool b200_supports_op(const b200_device_info&, const ggml_tensor* op) {
    const ggml_tensor* s0 = op->src[0];
    const ggml_tensor* s1 = op->src[1];
    switch (op->op) {
        case GGML_OP_NONE:
        case GGML_OP_RESHAPE:
        case GGML_OP_VIEW:
        case GGML_OP_PERMUTE:
        case GGML_OP_TRANSPOSE:
            return true;
        case GGML_OP_ADD:
        case GGML_OP_SUB:
        case GGML_OP_MUL:
        case GGML_OP_DIV: {
            int a = s0->type, b = s1->type, d = op->type;
            bool ok = (a == GGML_TYPE_F32 && b == GGML_TYPE_F32 && d == GGML_TYPE_F32) ||
                      (a == GGML_TYPE_F16 && b == GGML_TYPE_F16 && d == GGML_TYPE_F16) ||
                      (a == GGML_TYPE_F32 && b == GGML_TYPE_F16 && d == GGML_TYPE_F32) ||
                      (a == GGML_TYPE_F16 && b == GGML_TYPE_F32 && (d == GGML_TYPE_F32 || d == GGML_TYPE_F16)) ||
                      (a == GGML_TYPE_BF16 && b == GGML_TYPE_BF16 && d == GGML_TYPE_BF16);
            return ok;
        }
        case GGML_OP_UNARY:
            switch (ggml_get_unary_op(op)) {
                case GGML_UNARY_OP_ABS: case GGML_UNARY_OP_SGN: case GGML_UNARY_OP_NEG: case GGML_UNARY_OP_STEP: case GGML_UNARY_OP_TANH:
                case GGML_UNARY_OP_ELU: case GGML_UNARY_OP_RELU: case GGML_UNARY_OP_SIGMOID: case GGML_UNARY_OP_GELU: case GGML_UNARY_OP_GELU_QUICK:
                case GGML_UNARY_OP_SILU: case GGML_UNARY_OP_HARDSWISH: case GGML_UNARY_OP_HARDSIGMOID: case GGML_UNARY_OP_EXP: case GGML_UNARY_OP_EXPM1:
                case GGML_UNARY_OP_SOFTPLUS: case GGML_UNARY_OP_GELU_ERF: case GGML_UNARY_OP_FLOOR: case GGML_UNARY_OP_CEIL: case GGML_UNARY_OP_ROUND:
                case GGML_UNARY_OP_TRUNC:
                    return (s0->type == GGML_TYPE_F32 && op->type == GGML_TYPE_F32) || (s0->type == GGML_TYPE_F16 && op->type == GGML_TYPE_F16);
                default:
                    return false;
            }
        case GGML_OP_SCALE:
        case GGML_OP_CLAMP:
        case GGML_OP_SQR:
        case GGML_OP_SQRT:
        case GGML_OP_SIN:
        case GGML_OP_COS:
        case GGML_OP_LOG:
        case GGML_OP_LEAKY_RELU:
            return (s0->type == GGML_TYPE_F32 && op->type == GGML_TYPE_F32) || (s0->type == GGML_TYPE_F16 && op->type == GGML_TYPE_F16);
        case GGML_OP_GLU:
            return s0->type == GGML_TYPE_F32 && op->type == GGML_TYPE_F32 && ggml_is_contiguous_1(s0) && (!s1 || (s1->type == GGML_TYPE_F32 && ggml_is_contiguous_1(s1))) &&
                   ggml_get_glu_op(op) <= GGML_GLU_OP_GEGLU_QUICK && ggml_get_glu_op(op) != GGML_GLU_OP_SWIGLU_OAI;
        case GGML_OP_CPY:
        case GGML_OP_DUP:
        case GGML_OP_CONT: {
            const ggml_tensor* d = op->op == GGML_OP_CPY ? s1 : op;
            int a = s0->type, b = d->type;
            if (a == b) return ggml_type_size((ggml_type)a) <= 8 && ggml_blck_size((ggml_type)a) == 1;
            return is_fp(s0) && (b == GGML_TYPE_F32 || b == GGML_TYPE_F16 || b == GGML_TYPE_BF16);
        }
        case GGML_OP_CONCAT:
            return s0->type == s1->type && s0->type == op->type && (ggml_type_size(op->type) == 4 || ggml_type_size(op->type) == 2) && ggml_blck_size(op->type) == 1;
        case GGML_OP_REPEAT:
            return s0->type == op->type && (ggml_type_size(op->type) == 4 || ggml_type_size(op->type) == 2) && ggml_blck_size(op->type) == 1;
        case GGML_OP_PAD:
            return is_f32(s0) && is_f32(op);
        case GGML_OP_UPSCALE: {
            int mode = ggml_get_op_params_i32(op, 0);
            int m = mode & 0xFF;
            if (mode & GGML_SCALE_FLAG_ANTIALIAS) return false;
            return is_f32(s0) && is_f32(op) && (m == GGML_SCALE_MODE_NEAREST || m == GGML_SCALE_MODE_BILINEAR);
        }
        case GGML_OP_TIMESTEP_EMBEDDING:
            return is_f32(s0) && is_f32(op) && ggml_is_contiguous(s0);
        case GGML_OP_GET_ROWS:
            return is_fp(s0) && s1->type == GGML_TYPE_I32 && is_f32(op);
        case GGML_OP_ARANGE:
        case GGML_OP_FILL:
            return is_f32(op) && ggml_is_contiguous(op);
        case GGML_OP_SUM_ROWS:
        case GGML_OP_MEAN:
            return is_f32(s0) && is_f32(op);
        case GGML_OP_GROUP_NORM:
            return is_f32(s0) && is_f32(op) && ggml_is_contiguous(s0) && ggml_is_contiguous(op);
        case GGML_OP_NORM:
        case GGML_OP_RMS_NORM:
        case GGML_OP_L2_NORM:
            return is_f32(s0) && is_f32(op) && s0->nb[0] == 4 && op->nb[0] == 4;
        case GGML_OP_SOFT_MAX: {
            if (!is_f32(s0) || !is_f32(op) || !ggml_is_contiguous(op) || s0->nb[0] != 4) return false;
            if (op->src[2]) return false;   // attention sinks: not emitted by the diffusion graphs
            if (s1 && !(s1->type == GGML_TYPE_F32 || s1->type == GGML_TYPE_F16)) return false;
            if (s1 && !ggml_is_contiguous(s1)) return false;
            return s0->ne[0] * 4 <= 200 * 1024;
        }
        case GGML_OP_IM2COL:
            return (s1->type == GGML_TYPE_F32 || s1->type == GGML_TYPE_F16) && (op->type == GGML_TYPE_F16 || op->type == GGML_TYPE_F32) &&
                   s1->nb[0] == ggml_type_size(s1->type) && ggml_is_contiguous(op);
        case GGML_OP_MUL_MAT: {
            if (!is_f32(op) || !ggml_is_contiguous(op)) return false;
            if (!(s0->type == GGML_TYPE_F32 || s0->type == GGML_TYPE_F16 || s0->type == GGML_TYPE_BF16)) return false;
            if (!is_fp(s1)) return false;
            if (s1->ne[2] % s0->ne[2] || s1->ne[3] % s0->ne[3]) return false;
            return true;
        }
        case GGML_OP_FLASH_ATTN_EXT: {
            const ggml_tensor* q = s0; const ggml_tensor* k = s1; const ggml_tensor* v = op->src[2];
            const ggml_tensor* mask = op->src[3];
            if (op->src[4]) return false;   // sinks
            float max_bias, softcap;
            memcpy(&max_bias, (const float*)op->op_params + 1, 4);
            memcpy(&softcap, (const float*)op->op_params + 2, 4);
            if (softcap != 0.0f) return false;
            if (!is_f32(q) || !is_f32(op)) return false;
            if (!(k->type == GGML_TYPE_F16 || k->type == GGML_TYPE_BF16) || v->type != k->type) return false;
            if (mask && mask->type != GGML_TYPE_F16) return false;
            if (q->ne[2] % k->ne[2]) return false;
            if (k->ne[1] * 4 > 200 * 1024) return false;
            if (k->ne[2] != v->ne[2]) return false;
            return ggml_is_contiguous(op);
        }
        default:
            return false;
    }
}

// ------------------------------------------------------------------------------------------------
// node dispatch
// ------------------------------------------------------------------------------------------------
static int run_node(b200_context* ctx, ggml_tensor* t) {
    cudaStream_t s = ctx->stream;
    const ggml_tensor* s0 = t->src[0];
    const ggml_tensor* s1 = t->src[1];
    switch (t->op) {
        case GGML_OP_ADD: return b200_launch_binary(s, B200_ADD, b200_make_td(s0), b200_make_td(s1), b200_make_td(t));
        case GGML_OP_SUB: return b200_launch_binary(s, B200_SUB, b200_make_td(s0), b200_make_td(s1), b200_make_td(t));
        case GGML_OP_MUL: return b200_launch_binary(s, B200_MUL, b200_make_td(s0), b200_make_td(s1), b200_make_td(t));
        case GGML_OP_DIV: return b200_launch_binary(s, B200_DIV, b200_make_td(s0), b200_make_td(s1), b200_make_td(t));
        case GGML_OP_UNARY: return b200_launch_unary(s, (int)ggml_get_unary_op(t), b200_make_td(s0), b200_make_td(t));
        case GGML_OP_SCALE: {
            float sc, bi;
            memcpy(&sc, (const float*)t->op_params + 0, 4);
            memcpy(&bi, (const float*)t->op_params + 1, 4);
            return b200_launch_scalar_op(s, B200_SCALE, b200_make_td(s0), b200_make_td(t), sc, bi);
        }
        case GGML_OP_CLAMP: {
            float lo, hi;
            memcpy(&lo, (const float*)t->op_params + 0, 4);
            memcpy(&hi, (const float*)t->op_params + 1, 4);
            return b200_launch_scalar_op(s, B200_CLAMP, b200_make_td(s0), b200_make_td(t), lo, hi);
        }
        case GGML_OP_SQR: return b200_launch_scalar_op(s, B200_SQR, b200_make_td(s0), b200_make_td(t), 0, 0);
        case GGML_OP_SQRT: return b200_launch_scalar_op(s, B200_SQRT, b200_make_td(s0), b200_make_td(t), 0, 0);
        case GGML_OP_SIN: return b200_launch_scalar_op(s, B200_SIN, b200_make_td(s0), b200_make_td(t), 0, 0);
        case GGML_OP_COS: return b200_launch_scalar_op(s, B200_COS, b200_make_td(s0), b200_make_td(t), 0, 0);
        case GGML_OP_LOG: return b200_launch_scalar_op(s, B200_LOG, b200_make_td(s0), b200_make_td(t), 0, 0);
        case GGML_OP_LEAKY_RELU: {
            float slope;
            memcpy(&slope, t->op_params, 4);
            return b200_launch_scalar_op(s, B200_LEAKY_RELU, b200_make_td(s0), b200_make_td(t), slope, 0);
        }
        case GGML_OP_GLU: {
            b200_td b;
            if (s1) b = b200_make_td(s1);
            return b200_launch_glu(s, (int)ggml_get_glu_op(t), b200_make_td(s0), s1 ? &b : nullptr, b200_make_td(t), ggml_get_op_params_i32(t, 1) != 0);
        }
        case GGML_OP_CPY: return b200_launch_copy(s, b200_make_td(s0), b200_make_td(s1));
        case GGML_OP_DUP:
        case GGML_OP_CONT: return b200_launch_copy(s, b200_make_td(s0), b200_make_td(t));
        case GGML_OP_CONCAT: return b200_launch_concat(s, b200_make_td(s0), b200_make_td(s1), b200_make_td(t), ggml_get_op_params_i32(t, 0));
        case GGML_OP_REPEAT: return b200_launch_repeat(s, b200_make_td(s0), b200_make_td(t));
        case GGML_OP_PAD: return b200_launch_pad(s, b200_make_td(s0), b200_make_td(t), t->op_params, ggml_get_op_params_i32(t, 8) != 0);
        case GGML_OP_UPSCALE: return b200_launch_upscale(s, b200_make_td(s0), b200_make_td(t), ggml_get_op_params_i32(t, 0));
        case GGML_OP_TIMESTEP_EMBEDDING:
            return b200_launch_timestep_embedding(s, b200_make_td(s0), b200_make_td(t), ggml_get_op_params_i32(t, 0), ggml_get_op_params_i32(t, 1));
        case GGML_OP_GET_ROWS: return b200_launch_get_rows(s, b200_make_td(s0), b200_make_td(s1), b200_make_td(t));
        case GGML_OP_ARANGE: {
            float start, step;
            memcpy(&start, (const float*)t->op_params + 0, 4);
            memcpy(&step, (const float*)t->op_params + 2, 4);
            return b200_launch_arange(s, b200_make_td(t), start, step);
        }
        case GGML_OP_FILL: {
            float c;
            memcpy(&c, t->op_params, 4);
            return b200_launch_fill(s, b200_make_td(t), c);
        }
        case GGML_OP_SUM_ROWS: return b200_launch_sum_rows(s, b200_make_td(s0), b200_make_td(t), false);
        case GGML_OP_MEAN: return b200_launch_sum_rows(s, b200_make_td(s0), b200_make_td(t), true);
        case GGML_OP_GROUP_NORM: {
            float eps;
            memcpy(&eps, (const float*)t->op_params + 1, 4);
            return b200_launch_group_norm(s, b200_make_td(s0), b200_make_td(t), ggml_get_op_params_i32(t, 0), eps);
        }
        case GGML_OP_NORM:
        case GGML_OP_RMS_NORM:
        case GGML_OP_L2_NORM: {
            float eps;
            memcpy(&eps, t->op_params, 4);
            int kind = t->op == GGML_OP_NORM ? B200_NORM_LAYER : (t->op == GGML_OP_RMS_NORM ? B200_NORM_RMS : B200_NORM_L2);
            return b200_launch_norm(s, kind, b200_make_td(s0), b200_make_td(t), eps);
        }
        case GGML_OP_SOFT_MAX: {
            float scale, max_bias;
            memcpy(&scale, (const float*)t->op_params + 0, 4);
            memcpy(&max_bias, (const float*)t->op_params + 1, 4);
            b200_td m;
            if (s1) m = b200_make_td(s1);
            return b200_launch_soft_max(s, b200_make_td(s0), s1 ? &m : nullptr, b200_make_td(t), scale, max_bias);
        }
        case GGML_OP_IM2COL: {
            const int32_t* p = t->op_params;
            return b200_launch_im2col(s, b200_make_td(s1), b200_make_td(t), s0->ne[0], s0->ne[1], p[0], p[1], p[2], p[3], p[4], p[5], p[6] == 1);
        }
        case GGML_OP_MUL_MAT: return op_mul_mat(ctx, t);
        case GGML_OP_FLASH_ATTN_EXT: return op_flash_attn(ctx, t);
        default: return -1;
    }
}

static inline bool node_is_noop(const ggml_tensor* t) {
    return t->op == GGML_OP_NONE || t->op == GGML_OP_RESHAPE || t->op == GGML_OP_VIEW || t->op == GGML_OP_PERMUTE || t->op == GGML_OP_TRANSPOSE ||
           ggml_is_empty(t);
}

enum ggml_status b200_graph_compute(b200_context* ctx, ggml_cgraph* cgraph) {
    B200_CUDA_CHECK(cudaSetDevice(ctx->device));
    // the previous graph's device time is finalised here (the host has synchronised in between: it read the result)
    if ((ctx->opt_timing && ctx->timing_pending) || !ctx->kt_pending.empty()) b200_context_finalize_timing(ctx);
    ws_begin_graph(ctx);
    const bool timing = ctx->opt_timing && !ctx->timing_pending;
    if (timing) cudaEventRecord(ctx->ev_start, ctx->stream);

    for (int i = 0; i < cgraph->n_nodes; ++i) {
        ggml_tensor* t = cgraph->nodes[i];
        if (node_is_noop(t)) continue;
        if ((t->flags & GGML_TENSOR_FLAG_COMPUTE) == 0) continue;
        int n = run_node(ctx, t);
        if (n < 0) {
            GGML_LOG_ERROR("ggml-b200: node %d (%s, %s) is not executable on this backend\n", i, ggml_op_name(t->op), t->name);
            return GGML_STATUS_FAILED;
        }
        ctx->stats.kernel_launches += (uint64_t)n;
        ctx->stats.nodes_executed++;
#ifdef B200_DEBUG_SYNC
        {
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) {
                fprintf(stderr, "[ggml-b200] node %d %s failed: %s\n", i, ggml_op_name(t->op), cudaGetErrorString(e));
                return GGML_STATUS_FAILED;
            }
        }
#endif
    }
    if (timing) {
        cudaEventRecord(ctx->ev_stop, ctx->stream);
        ctx->timing_pending = true;
    }
    ctx->stats.graphs++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        GGML_LOG_ERROR("ggml-b200: launch error in graph_compute: %s\n", cudaGetErrorString(e));
        return GGML_STATUS_FAILED;
    }
    return GGML_STATUS_SUCCESS;
}
