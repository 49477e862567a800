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
#include <atomic>
#include <chrono>
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
    for (int k = 0; k < 3; ++k)
        if (cudaStreamCreateWithFlags(&ctx->side[k], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); ctx->side[k] = nullptr; }
    cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming);
    cudaEventCreate(&ctx->ev_start);
    cudaEventCreate(&ctx->ev_stop);
    if (cudaMalloc(&ctx->gn_counters, B200_GN_COUNTERS * sizeof(unsigned)) == cudaSuccess) cudaMemset(ctx->gn_counters, 0, B200_GN_COUNTERS * sizeof(unsigned));
    else { cudaGetLastError(); ctx->gn_counters = nullptr; }
    ctx->opt_fusion = env_flag("GGML_B200_FUSION", 1) != 0;
    ctx->opt_tc_gemm = env_flag("GGML_B200_TC_GEMM", 1) != 0;
    ctx->opt_timing = env_flag("GGML_B200_TIMING", 1) != 0;
    ctx->opt_cuda_graphs = env_flag("GGML_B200_CUDA_GRAPHS", 1) != 0;
    ctx->opt_fused_attn = env_flag("GGML_B200_FUSED_ATTN", 1) != 0;
    ctx->opt_implicit_conv = env_flag("GGML_B200_IMPLICIT_CONV", 1) != 0;
    ctx->opt_early_weights = env_flag("GGML_B200_EARLY_WEIGHTS", 0) != 0;   // measured neutral on the SD1.5 step (profiles/r01_summary.md): off by default
    ctx->opt_chain_fusion = env_flag("GGML_B200_CHAIN_FUSION", 1) != 0;
    ctx->opt_gemv = env_flag("GGML_B200_GEMV", 1) != 0;
    ctx->opt_precise_f32 = env_flag("GGML_B200_PRECISE_F32", 1) != 0;
    ctx->opt_q8_activations = env_flag("GGML_B200_Q8_ACT", 1) != 0;
    ctx->opt_side_streams = env_flag("GGML_B200_SIDE_STREAMS", 1) != 0;
    ctx->opt_wprefetch = env_flag("GGML_B200_WPREFETCH", 0) != 0;
    ctx->opt_fold_batch = env_flag("GGML_B200_FOLD_BATCH", 0) != 0;   // written at the end of round 1, not yet measured: off by default
    return ctx;
}

b200_context::~b200_context() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (int k = 0; k < 3; ++k) if (side[k]) { cudaStreamSynchronize(side[k]); cudaStreamDestroy(side[k]); }
    for (auto e : ev_pool) cudaEventDestroy(e);
    for (auto& c : ws.chunks) cudaFree(c.base);
    if (gn_counters) cudaFree(gn_counters);
    b200_peer_close(this);
    for (auto& kv : plans) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    for (auto& p : kt_pending) { cudaEventDestroy(p.start); cudaEventDestroy(p.stop); }
    for (auto e : kt_free) cudaEventDestroy(e);
    if (copy_event) cudaEventDestroy(copy_event);
    if (ev_start) cudaEventDestroy(ev_start);
    if (ev_stop) cudaEventDestroy(ev_stop);
    if (stream) cudaStreamDestroy(stream);
}

static void drop_cuda_graphs(b200_context* ctx);
static void b200_debug_refuse_op(int op);

int b200_context_set_option(b200_context* ctx, const char* key, int value) {
    // captured plans embody the options they were recorded under
    if (strcmp(key, "timing") && strcmp(key, "kernel_timing")) { drop_cuda_graphs(ctx); ctx->plans.clear(); }
    if (!strcmp(key, "fusion")) ctx->opt_fusion = value != 0;
    else if (!strcmp(key, "tc_gemm")) ctx->opt_tc_gemm = value != 0;
    else if (!strcmp(key, "timing")) ctx->opt_timing = value != 0;
    else if (!strcmp(key, "cuda_graphs")) ctx->opt_cuda_graphs = value != 0;
    else if (!strcmp(key, "kernel_timing")) ctx->opt_kernel_timing = value != 0;
    else if (!strcmp(key, "fused_attn")) ctx->opt_fused_attn = value != 0;
    else if (!strcmp(key, "side_streams")) ctx->opt_side_streams = value != 0;
    else if (!strcmp(key, "wprefetch")) ctx->opt_wprefetch = value != 0;
    else if (!strcmp(key, "implicit_conv")) ctx->opt_implicit_conv = value != 0;
    else if (!strcmp(key, "early_weights")) ctx->opt_early_weights = value != 0;
    else if (!strcmp(key, "chain_fusion")) ctx->opt_chain_fusion = value != 0;
    else if (!strcmp(key, "gemv")) ctx->opt_gemv = value != 0;
    else if (!strcmp(key, "fold_batch")) ctx->opt_fold_batch = value != 0;
    else if (!strcmp(key, "precise_f32")) ctx->opt_precise_f32 = value != 0;
    else if (!strcmp(key, "q8_activations")) ctx->opt_q8_activations = value != 0;
    else if (!strcmp(key, "debug_refuse_op")) b200_debug_refuse_op(value);
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
    if (ctx->capturing) {           // allocation is illegal inside stream capture: abandon the capture, caller re-runs eagerly
        ctx->capture_overflow = true;
        return nullptr;
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

using operand = b200_operand;

// Bring `t` ([K, rows, b2, b3]) into a form the tcgen05 GEMM can read as type `want`: in place when possible,
// otherwise packed (converted, K padded to 16 bytes) into workspace.  Returns false on allocation failure.
static const void* get_dequantised_weight(b200_context* ctx, const ggml_tensor* w, int* launches);

// events for the fork / join edges between the main stream and the side streams; none is created while a capture is running
static cudaEvent_t side_event(b200_context* ctx) {
    if (ctx->ev_next < ctx->ev_pool.size()) return ctx->ev_pool[ctx->ev_next++];
    if (ctx->capturing) return nullptr;
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->ev_pool.push_back(e);
    ctx->ev_next++;
    return e;
}

static bool prepare_operand(b200_context* ctx, const ggml_tensor* t, int want, operand* out, int* launches) {
    if (t->type == GGML_TYPE_Q8_0) {
        if (want != GGML_TYPE_F16) return false;
        const void* p = get_dequantised_weight(ctx, t, launches);
        if (!p) return false;
        *out = operand{p, GGML_TYPE_F16, t->ne[0], t->ne[0] * t->ne[1], t->ne[0] * t->ne[1] * t->ne[2]};
        return true;
    }
    if ((int)t->type == want && tma_compatible(t)) {
        const int64_t es = fp_size(want);
        *out = operand{t->data, want, (int64_t)t->nb[1] / es, (int64_t)t->nb[2] / es, (int64_t)t->nb[3] / es};
        return true;
    }
    // the same tensor node feeding several contractions (x_norm -> to_q / to_k / to_v, the text context -> 32 k/v projections) is packed
    // once per graph execution: a ggml tensor node is a value, its buffer is not overwritten before its last consumer has run
    auto hit = ctx->pack_cache.find(std::make_pair(t, want));
    if (hit != ctx->pack_cache.end()) {
        if (!ctx->pack_origins.empty()) {      // packed on another stream: this stream waits for that pack kernel
            auto o = ctx->pack_origins.find(std::make_pair(t, want));
            if (o != ctx->pack_origins.end() && o->second.stream != ctx->stream) cudaStreamWaitEvent(ctx->stream, o->second.ready, 0);
        }
        *out = hit->second;
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
    if (ctx->on_side) {
        cudaEvent_t e = side_event(ctx);
        if (!e) return true;                   // (no event left inside a capture: usable by this launch, not shared)
        cudaEventRecord(e, ctx->stream);
        ctx->pack_origins[std::make_pair(t, want)] = b200_context::pack_origin{ctx->stream, e};
    }
    ctx->pack_cache[std::make_pair(t, want)] = *out;
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
    if (n == 2) { n = 1; ctx->stats.ext[5] += 1; }   // CTA-pair kernel (gemm_tc2.cu)
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
    if (src0->type == GGML_TYPE_F16 || src0->type == GGML_TYPE_Q8_0) return GGML_TYPE_F16;
    if (src0->type == GGML_TYPE_BF16) return GGML_TYPE_BF16;
    return GGML_TYPE_F32;
}

// epilogue work absorbed from the nodes that follow a MUL_MAT in the graph (see try_fuse_mul_mat)
struct mm_fusion {
    float* out = nullptr;         // write here instead of dst->data (same [M, N, b2, b3] layout)
    const float* bias = nullptr;  // f32 vector
    int bias_mode = 0;            // 1: per output row m (Linear bias), 2: per n (conv bias: n == output channel)
    const float* residual = nullptr;   // same [M, N] layout as out, added last
    const float* gate = nullptr;       // f32 [M]: value = residual + gate[m] * value (set together with residual)
    const ggml_tensor* src1_pre = nullptr;   // activation to read instead of src[1] (same shape): the input of a unary op folded in
    int pre_act = 0;                          // 1: SiLU applied to src1_pre on load (only the few-row GEMV path can do this)
    int act = 0;                              // activation after the bias, before the residual: 1 SiLU, 2 GELU (tanh form) -- the epilogue's act_fn
    void* d16 = nullptr;                      // 16-bit copy of the result for the contraction that consumes it (b200_gemm_args::D16)
    int d16_type = 0;
    bool skip_f32 = false;
    int* d16_done = nullptr;
    bool d16_strict = false;                  // decline (-2, nothing written) rather than run without the 16-bit copy
    int64_t geglu = 0;                        // > 0: GEGLU projection (b200_gemm_args::geglu): d16 = x * gelu(gate) [tokens][geglu] is the only output
};

static int op_mul_mat(b200_context* ctx, ggml_tensor* dst, const mm_fusion* fz = nullptr) {
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

    // a handful of activation rows against in-place F16/BF16 weights (embedding MLPs): weight-streaming GEMV, no operand packing
    if (!(fz && (fz->act || fz->gate || fz->d16_strict || fz->geglu))) {
        const ggml_tensor* x = fz && fz->src1_pre ? fz->src1_pre : src1;
        if (ctx->opt_gemv && ne02 * ne03 * ne12 * ne13 == 1 && N <= 4 && (src0->type == GGML_TYPE_F16 || src0->type == GGML_TYPE_BF16) &&
            rows_unit_stride(src0) && x->type == GGML_TYPE_F32 && x->nb[0] == 4 && x->nb[1] % 4 == 0 && dst->nb[0] == 4 &&
            b200_gemv_supported((int)src0->type, M, N, K, src0->data, (int64_t)src0->nb[1] / 2, x->data)) {
            float* out = fz && fz->out ? fz->out : (float*)dst->data;
            const float* bias = fz && fz->bias && fz->bias_mode == 1 ? fz->bias : nullptr;
            if (!(fz && fz->bias && fz->bias_mode != 1)) {
                int n = b200_launch_gemv(ctx->stream, (int)src0->type, src0->data, (int64_t)src0->nb[1] / 2, (const float*)x->data, (int64_t)x->nb[1] / 4, out,
                                         (int64_t)dst->nb[1] / 4, M, N, K, bias, fz ? fz->residual : nullptr, (int64_t)dst->nb[1] / 4, fz ? fz->pre_act : 0);
                if (n > 0) { ctx->stats.reserved[6] += 1; return n; }
            }
        }
        if (fz && fz->src1_pre) return -2;     // only the GEMV can fold the unary op
    }

    if (fz && fz->geglu && (ct == GGML_TYPE_F32 || !ctx->opt_tc_gemm)) return -2;
    if (ct == GGML_TYPE_F32 && ctx->opt_precise_f32 && ctx->opt_tc_gemm && !(fz && (fz->act || fz->gate))) {
        // F32 x F32 (attention GEMMs of the reference's default graph, F32 Linear weights): the CPU oracle computes true f32 dot products
        // (ggml-cpu.c:1406 with vec_dot_f32); a single TF32 pass would keep 10 mantissa bits of each operand.  3xTF32: x = hi + lo with hi
        // exactly representable in TF32; D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, three tensor-core passes chained through the residual
        // input of the epilogue, small terms first.  Error ~2^-21 relative: f32 class.
        auto split = [&](const ggml_tensor* t, operand* hi, operand* lo) -> bool {
            const int key_hi = 1000, key_lo = 1001;
            auto h = ctx->pack_cache.find(std::make_pair(t, key_hi));
            auto l = ctx->pack_cache.find(std::make_pair(t, key_lo));
            if (h != ctx->pack_cache.end() && l != ctx->pack_cache.end()) { *hi = h->second; *lo = l->second; return true; }
            const int64_t kpad = (t->ne[0] + 3) / 4 * 4;
            const int64_t rows = t->ne[1] * t->ne[2] * t->ne[3];
            float* bh = (float*)ws_alloc(ctx, (size_t)(rows * kpad * 4));
            float* bl = (float*)ws_alloc(ctx, (size_t)(rows * kpad * 4));
            if (!bh || !bl) return false;
            int n = b200_launch_split_tf32(ctx->stream, b200_make_td(t), bh, bl, kpad);
            if (n < 0) return false;
            launches += n;
            *hi = operand{bh, GGML_TYPE_F32, kpad, kpad * t->ne[1], kpad * t->ne[1] * t->ne[2]};
            *lo = operand{bl, GGML_TYPE_F32, kpad, kpad * t->ne[1], kpad * t->ne[1] * t->ne[2]};
            ctx->pack_cache[std::make_pair(t, key_hi)] = *hi;
            ctx->pack_cache[std::make_pair(t, key_lo)] = *lo;
            return true;
        };
        operand ah, al, bh, bl;
        if (src0->type == GGML_TYPE_F32 && src1->type == GGML_TYPE_F32 && split(src0, &ah, &al) && split(src1, &bh, &bl)) {
            bool ok = true;
            for (int64_t i3 = 0; i3 < ne13 && ok; ++i3) {
                char* out_base = fz && fz->out ? (char*)fz->out : (char*)dst->data;
                float* D = (float*)(out_base + i3 * dst->nb[3]);
                const operand* As[3] = {&al, &ah, &ah};
                const operand* Bs[3] = {&bh, &bl, &bh};
                for (int pass = 0; pass < 3 && ok; ++pass) {
                    b200_gemm_args g;
                    memset(&g, 0, sizeof(g));
                    g.A = (const char*)As[pass]->ptr + (i3 / r3) * As[pass]->b3_stride * 4;
                    g.B = (const char*)Bs[pass]->ptr + i3 * Bs[pass]->b3_stride * 4;
                    g.type = GGML_TYPE_F32;
                    g.M = M; g.N = N; g.K = K;
                    g.lda = As[pass]->ld; g.ldb = Bs[pass]->ld;
                    g.batch = ne12;
                    g.a_batch_stride = As[pass]->batch_stride;
                    g.b_batch_stride = Bs[pass]->batch_stride;
                    g.a_bcast = r2;
                    g.D = D;
                    g.ldd = dst->nb[1] / 4;
                    g.d_batch_stride = dst->nb[2] / 4;
                    g.ldr = g.ldd;
                    if (pass == 0) { if (fz && fz->residual) g.residual = fz->residual + i3 * (dst->nb[3] / 4); }
                    else g.residual = D;                                     // accumulate onto the previous pass (same element, same thread)
                    if (pass == 2 && fz && fz->bias) { g.bias = fz->bias; g.bias_mode = fz->bias_mode; }
                    const int n = launch_tc(ctx, g);
                    if (n < 0) ok = false;
                    else launches += n;
                }
            }
            if (ok) return launches;
            return -1;       // a half-executed chain must not be redone differently: report the failure
        }
    }

    operand a, b;
    if (!prepare_operand(ctx, src0, ct, &a, &launches)) return -1;
    bool have_b = false;
    if (src0->type == GGML_TYPE_Q8_0 && src1->type == GGML_TYPE_F32 && ctx->opt_q8_activations && K % 32 == 0) {
        // the oracle quantises the activation rows to Q8_0 as well (q8_0 x q8_0 dot, ggml-cpu.c:1480-1510): contract the same values
        const int key = 2000;
        auto hit = ctx->pack_cache.find(std::make_pair(src1, key));
        if (hit != ctx->pack_cache.end()) { b = hit->second; have_b = true; }
        else {
            const int64_t rows = src1->ne[1] * src1->ne[2] * src1->ne[3];
            void* buf = ws_alloc(ctx, (size_t)(rows * K * 2));
            if (!buf) return -1;
            const int n = b200_launch_pack_rows_q8_roundtrip(ctx->stream, b200_make_td(src1), buf, K);
            if (n > 0) {
                launches += n;
                b = operand{buf, GGML_TYPE_F16, K, K * src1->ne[1], K * src1->ne[1] * src1->ne[2]};
                ctx->pack_cache[std::make_pair(src1, key)] = b;
                have_b = true;
            }
        }
    }
    if (!have_b && !prepare_operand(ctx, src1, ct, &b, &launches)) return -1;

    // One weight matrix against a batch of activation matrices that lie back to back in memory (a Linear on [C, L, N] tokens of a
    // batched-CFG graph): fold the batch into the N dimension -- fewer, fuller tiles (N = 64 + 64 fills one 128-row tile instead of
    // two half-empty ones) and one pass over the weights instead of one per image.
    const bool fold = ctx->opt_fold_batch && ne02 == 1 && ne03 == 1 && ne12 * ne13 > 1 && b.batch_stride == b.ld * N &&
                      (ne13 == 1 || b.b3_stride == b.batch_stride * ne12) && dst->nb[2] == dst->nb[1] * (size_t)N &&
                      (ne13 == 1 || dst->nb[3] == dst->nb[2] * (size_t)ne12) && !(fz && fz->bias && fz->bias_mode == 2);
    const int64_t Ng = fold ? N * ne12 * ne13 : N;            // rows of the activation operand per GEMM
    const int64_t nb12 = fold ? 1 : ne12, nb13 = fold ? 1 : ne13, rr2 = fold ? 1 : r2, rr3 = fold ? 1 : r3;

    for (int64_t i3 = 0; i3 < nb13; ++i3) {
        const int64_t es = fp_size(ct);
        b200_gemm_args g;
        memset(&g, 0, sizeof(g));
        g.A = (const char*)a.ptr + (i3 / rr3) * a.b3_stride * es;
        g.B = (const char*)b.ptr + i3 * b.b3_stride * es;
        g.type = ct;
        g.M = M; g.N = Ng; g.K = K;
        g.lda = a.ld; g.ldb = b.ld;
        g.batch = nb12;
        g.a_batch_stride = a.batch_stride;
        g.b_batch_stride = b.batch_stride;
        g.a_bcast = rr2;
        char* out_base = fz && fz->out ? (char*)fz->out : (char*)dst->data;
        g.D = (float*)(out_base + i3 * dst->nb[3]);
        g.ldd = dst->nb[1] / 4;
        g.d_batch_stride = dst->nb[2] / 4;
        if (fz && fz->bias) { g.bias = fz->bias; g.bias_mode = fz->bias_mode; }
        if (fz && fz->residual) { g.residual = fz->residual; g.ldr = g.ldd; }
        if (fz) { g.act = fz->act; g.gate = fz->gate; }
        if (fz && fz->d16 && nb13 == 1) { g.D16 = fz->d16; g.d16_type = fz->d16_type; g.skip_f32 = fz->skip_f32 ? 1 : 0; g.d16_done = fz->d16_done; g.d16_strict = fz->d16_strict ? 1 : 0; }
        else if (fz && (fz->d16_strict || fz->geglu)) return -2;
        if (fz && fz->geglu) g.geglu = fz->geglu;
        // model weights read in place are constants of the graph: their first ring-full may be fetched before the PDL wait.  Not for
        // the first kernel of a graph_compute (its stream predecessor belongs to an earlier call, e.g. a weight update).
        if (ctx->opt_early_weights && a.ptr == src0->data && src0->buffer && src0->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS &&
            (ctx->launched_any || launches > 0))
            g.early = 1;
        // (in place or a derived copy made at upload: constant either way, unless this very graph stores into weights -- LoRA apply)
        if (ctx->opt_wprefetch && !ctx->graph_writes_weights && src0->buffer && src0->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS &&
            (a.ptr == src0->data || src0->type == GGML_TYPE_Q8_0) && ne02 * ne03 == 1)
            g.wprefetch = 1;
        int n = launch_tc(ctx, g);
        if (n < 0) {
            if (fz && (fz->d16_strict || fz->geglu)) return -2;   // declined before anything was launched: the caller runs the node in its turn / unfused
            if (fz && (fz->act || fz->gate)) return -1;      // the reference kernel has no activation / gate epilogue: fail loudly rather than skip it
            // CUDA-core reference kernel (debug option, or shapes the TMA cannot describe)
            n = 0;
            for (int64_t i2 = 0; i2 < nb12; ++i2) {
                int r = b200_launch_gemm_ref(ctx->stream, (const char*)g.A + (i2 / rr2) * a.batch_stride * es, ct, a.ld * es,
                                             (const char*)g.B + i2 * b.batch_stride * es, ct, b.ld * es, g.D + i2 * g.d_batch_stride, g.ldd, M, Ng, K);
                if (r < 0) return -1;
                n += r;
                ctx->stats.ext[0] += (uint64_t)r;      // CUDA-core reference GEMM: visible to the tests (must be 0 on model paths)
            }
            if (fz && fz->bias) {
                b200_td o;
                o.data = g.D; o.type = GGML_TYPE_F32;
                o.ne[0] = M; o.ne[1] = Ng; o.ne[2] = nb12; o.ne[3] = 1;
                o.nb[0] = 4; o.nb[1] = g.ldd * 4; o.nb[2] = g.d_batch_stride * 4; o.nb[3] = o.nb[2] * nb12;
                b200_td bv = o;
                bv.data = (void*)fz->bias;
                bv.ne[0] = fz->bias_mode == 1 ? M : 1; bv.ne[1] = fz->bias_mode == 2 ? Ng : 1; bv.ne[2] = 1; bv.ne[3] = 1;
                bv.nb[0] = 4; bv.nb[1] = 4; bv.nb[2] = bv.nb[3] = 4 * (fz->bias_mode == 1 ? M : Ng);
                int r = b200_launch_binary(ctx->stream, B200_ADD, o, bv, o);
                if (r < 0) return -1;
                n += r;
            }
            if (fz && fz->residual) {
                b200_td o;
                o.data = g.D; o.type = GGML_TYPE_F32;
                o.ne[0] = M; o.ne[1] = Ng; o.ne[2] = nb12; o.ne[3] = 1;
                o.nb[0] = 4; o.nb[1] = g.ldd * 4; o.nb[2] = g.d_batch_stride * 4; o.nb[3] = o.nb[2] * nb12;
                b200_td rv = o;
                rv.data = (void*)fz->residual;
                int r = b200_launch_binary(ctx->stream, B200_ADD, o, rv, o);
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
// graph-level help for the fused kernel (try_fuse_flash_attn): Q read through the strided view a skipped CONT would have copied,
// and an f16 copy of the result for the output projection that follows
struct fa_fusion {
    const b200_td* q_td = nullptr;
    int q_split_b = 0;                    // > 0: q_td is [d, Lq, H, q_split_b] (batch apart) instead of ggml's [d, Lq, H * B, 1]
    ggml_tensor* q_cont = nullptr;        // the skipped CONT: executed late if the fused kernel turns the shape down
    const ggml_tensor* shadow_act = nullptr;   // activation operand (f32 view of dst) of the next MUL_MAT
    // K / V read from the projection GEMM's f16 rows ([d, Lk, H, B] descriptors, see try_fuse_mul_mat); cast_dst: where the graph
    // expected the contiguous f16 tensor -- filled late if the fused kernel turns the shape down
    const b200_td* k_td = nullptr;
    const b200_td* v_td = nullptr;
    // the result in the layout of the CONT that follows ([d, H, Lq, B] contiguous) as f16 ONLY, for the output projection that is its
    // single reader (out_act: that projection's activation operand); set by try_fuse_flash_attn, honoured by the fused kernel or not at all
    const ggml_tensor* out_cont = nullptr;
    const ggml_tensor* out_act = nullptr;
    int out_h = 0, out_b = 0;
    bool out_used = false;
    // no CONT in between (batch 1: FLASH_ATTN_EXT -> views -> to_out): shadow_act is the projection's operand; out_skip: that projection
    // is the ONLY reader of the result, so the f32 tensor is not stored
    bool out_skip = false;
};

static int run_node(b200_context* ctx, ggml_tensor* t);

static int op_flash_attn(b200_context* ctx, ggml_tensor* dst, fa_fusion* fz = nullptr) {
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

    // Operands that never took ggml's [d, L, H * B] form (projection rows read in place, see fa_fusion): every tensor is described with
    // the batch as its own dimension, [d, L, H, B]; contiguous ones split trivially.  One fused launch; if the kernel turns it down the
    // skipped copies are produced after all and the generic paths below run.
    if (fz && (fz->k_td || fz->v_td || fz->q_split_b || fz->out_cont)) {
        int64_t B = 0;
        bool ok = ctx->opt_tc_gemm && ctx->opt_fused_attn && max_bias == 0.0f && !mask && NB == 1 && k->ne[3] == 1 && v->ne[3] == 1;
        auto want_b = [&](int64_t b) { if (b <= 0) return; if (B == 0) B = b; else if (B != b) ok = false; };
        if (fz->k_td) want_b(fz->k_td->ne[3]);
        if (fz->v_td) want_b(fz->v_td->ne[3]);
        if (fz->q_split_b) want_b(fz->q_split_b);
        if (fz->out_cont) want_b(fz->out_b);
        if (B == 0) B = 1;
        auto split = [&](b200_td td) {      // [x, L, H * B, 1] -> [x, L, H, B]
            if (td.ne[2] % B) { ok = false; return td; }
            td.ne[2] /= B; td.ne[3] = B; td.nb[3] = td.nb[2] * (size_t)td.ne[2];
            return td;
        };
        if (ok && (H % B || Hkv % B)) ok = false;
        if (ok) {
            const b200_td qd = fz->q_td ? (fz->q_split_b ? *fz->q_td : split(*fz->q_td)) : split(b200_make_td(q));
            const b200_td kd = fz->k_td ? *fz->k_td : split(b200_make_td(k));
            const b200_td vd = fz->v_td ? *fz->v_td : split(b200_make_td(v));
            b200_td dd = b200_make_td(dst);            // [dv, H * B, Lq, 1]: head-and-batch index hb = b * H + h
            const int64_t Hs = H / B;
            void* shadow = nullptr;
            const ggml_tensor* shadow_for = nullptr;
            int skip_f32 = 0;
            if (fz->out_cont && fz->out_h == Hs) {
                // write the CONT's layout [dv, H, Lq, B] directly, f16 only
                dd.ne[1] = Hs; dd.ne[3] = B;
                dd.nb[1] = (size_t)dv * 4; dd.nb[2] = (size_t)dv * 4 * Hs; dd.nb[3] = (size_t)dv * 4 * Hs * Lq;
                shadow = ws_alloc(ctx, (size_t)ggml_nelements(dst) * 2);
                shadow_for = fz->out_act;
                skip_f32 = 1;
                if (!shadow) ok = false;
            } else {
                dd.ne[1] = Hs; dd.ne[3] = B; dd.nb[3] = dd.nb[1] * (size_t)Hs;
                if (fz->shadow_act) {
                    shadow = ws_alloc(ctx, (size_t)ggml_nelements(dst) * 2);
                    shadow_for = fz->shadow_act;
                    if (shadow && fz->out_skip) skip_f32 = 1;
                }
            }
            if (ok && kd.ne[2] == vd.ne[2] && kd.ne[3] == B && vd.ne[3] == B && qd.ne[3] == B) {
                int n = b200_launch_flash_attn_fused(ctx->stream, qd, kd, nullptr, 0, vd, nullptr, dd, scale, shadow, skip_f32);
                if (n > 0) {
                    ctx->stats.reserved[2] += (uint64_t)n;
                    if (shadow && shadow_for)
                        ctx->pack_cache[std::make_pair(shadow_for, (int)GGML_TYPE_F16)] =
                            operand{shadow, GGML_TYPE_F16, shadow_for->ne[0], shadow_for->ne[0] * shadow_for->ne[1], shadow_for->ne[0] * shadow_for->ne[1] * shadow_for->ne[2]};
                    if (fz->q_td) ctx->stats.reserved[5] += 1;
                    if (skip_f32) { fz->out_used = fz->out_cont != nullptr; ctx->stats.ext[14] += 1; }
                    return launches + n;
                }
            }
        }
        // turned down: materialise what the graph expected (the allocator reserved those tensors until this node)
        if (fz->k_td) {
            const int n = b200_launch_pack_rows(ctx->stream, *fz->k_td, k->data, GGML_TYPE_F16, k->ne[0]);
            if (n < 0) return -1;
            launches += n;
        }
        if (fz->v_td) {
            const int n = b200_launch_pack_rows(ctx->stream, *fz->v_td, v->data, GGML_TYPE_F16, v->ne[0]);
            if (n < 0) return -1;
            launches += n;
        }
        if (fz->q_split_b) {
            if (!fz->q_cont) return -1;
            const int n = run_node(ctx, fz->q_cont);
            if (n < 0) return -1;
            launches += n;
            fz->q_td = nullptr; fz->q_cont = nullptr; fz->q_split_b = 0;
        }
    }

    // fused single-kernel path first, reading V in place (MN-major operand of the P.V product): no V^T pass, no workspace
    static int fa_vmn = -1;
    if (fa_vmn < 0) { const char* e = getenv("GGML_B200_FA_VMN"); fa_vmn = (e && *e) ? atoi(e) : 1; }
    if (fa_vmn && ctx->opt_tc_gemm && ctx->opt_fused_attn && max_bias == 0.0f && ct == GGML_TYPE_F16 && k->type == GGML_TYPE_F16 && v->type == GGML_TYPE_F16) {
        b200_td mtd;
        if (mask) mtd = b200_make_td(mask);
        void* shadow = nullptr;
        if (fz && fz->shadow_act) shadow = ws_alloc(ctx, (size_t)ggml_nelements(dst) * 2);
        const int skip = (shadow && fz->out_skip) ? 1 : 0;
        int n = b200_launch_flash_attn_fused(ctx->stream, fz && fz->q_td ? *fz->q_td : b200_make_td(q), b200_make_td(k), nullptr, 0, b200_make_td(v),
                                             mask ? &mtd : nullptr, b200_make_td(dst), scale, shadow, skip);
        if (n > 0 && skip) ctx->stats.ext[14] += 1;
        if (n < 0 && shadow) {
            shadow = nullptr;
            n = b200_launch_flash_attn_fused(ctx->stream, fz && fz->q_td ? *fz->q_td : b200_make_td(q), b200_make_td(k), nullptr, 0, b200_make_td(v),
                                             mask ? &mtd : nullptr, b200_make_td(dst), scale, nullptr);
        }
        if (n > 0) {
            ctx->stats.reserved[2] += (uint64_t)n;   // fused attention launches
            if (shadow) {
                const ggml_tensor* a = fz->shadow_act;
                ctx->pack_cache[std::make_pair(a, (int)GGML_TYPE_F16)] = operand{shadow, GGML_TYPE_F16, a->ne[0], a->ne[0] * a->ne[1], a->ne[0] * a->ne[1] * a->ne[2]};
            }
            if (fz && fz->q_td) ctx->stats.reserved[5] += 1;   // Q read in place: CONT skipped
            return launches + n;
        }
    }

    // V^T: [Lk, dv, Hkv, NB] view of v, packed to ct with Lk padded (the GEMM + softmax + GEMM path and head sizes outside the fused envelope)
    ggml_tensor vt = *v;
    vt.ne[0] = v->ne[1]; vt.nb[0] = v->nb[1];
    vt.ne[1] = v->ne[0]; vt.nb[1] = v->nb[0];
    void* vbuf = ws_alloc(ctx, (size_t)(Lk_pad * dv * Hkv * NB * es));
    if (!vbuf) return -1;
    int n = -1;
    if (v->type == (ggml_type)ct && v->nb[0] == 2) n = b200_launch_transpose_f16(ctx->stream, b200_make_td(v), vbuf, Lk_pad);
    if (n < 0) n = b200_launch_pack_rows(ctx->stream, b200_make_td(&vt), vbuf, ct, Lk_pad);
    if (n < 0) return -1;
    launches += n;

    // fused single-kernel path (tcgen05 QK^T and PV, online softmax between them): f16 K/V, d == dv, d % 8 == 0, d <= 192.
    // The kernel converts Q from f32 itself (any row stride), so no operand is packed for it.
    if (ctx->opt_tc_gemm && ctx->opt_fused_attn && max_bias == 0.0f && ct == GGML_TYPE_F16 && k->type == GGML_TYPE_F16) {
        b200_td mtd;
        if (mask) mtd = b200_make_td(mask);
        void* shadow = nullptr;
        if (fz && fz->shadow_act) shadow = ws_alloc(ctx, (size_t)ggml_nelements(dst) * 2);
        n = b200_launch_flash_attn_fused(ctx->stream, fz && fz->q_td ? *fz->q_td : b200_make_td(q), b200_make_td(k), vbuf, Lk_pad, b200_make_td(v),
                                         mask ? &mtd : nullptr, b200_make_td(dst), scale, shadow);
        if (n < 0 && shadow) {
            shadow = nullptr;
            n = b200_launch_flash_attn_fused(ctx->stream, fz && fz->q_td ? *fz->q_td : b200_make_td(q), b200_make_td(k), vbuf, Lk_pad, b200_make_td(v),
                                             mask ? &mtd : nullptr, b200_make_td(dst), scale, nullptr);
        }
        if (n > 0) {
            ctx->stats.reserved[2] += (uint64_t)n;   // fused attention launches
            if (shadow) {
                const ggml_tensor* a = fz->shadow_act;
                ctx->pack_cache[std::make_pair(a, (int)GGML_TYPE_F16)] = operand{shadow, GGML_TYPE_F16, a->ne[0], a->ne[0] * a->ne[1], a->ne[0] * a->ne[1] * a->ne[2]};
            }
            if (fz && fz->q_td) ctx->stats.reserved[5] += 1;   // Q read in place: CONT skipped
            return launches + n;
        }
    }
    if (fz && fz->q_cont) {     // composite path reads the contiguous Q: produce it now (its source is still intact, see try_skip_q_cont)
        n = run_node(ctx, fz->q_cont);
        if (n < 0) return -1;
        launches += n;
    }
    ctx->stats.ext[7] += 1;     // unfused attention (scores materialised)
    operand qa, ka;
    if (!prepare_operand(ctx, q, ct, &qa, &launches)) return -1;
    if (!prepare_operand(ctx, k, ct, &ka, &launches)) return -1;
    float* sbuf = (float*)ws_alloc(ctx, (size_t)(Lk * Lq * H * sizeof(float)));
    void* pbuf = ws_alloc(ctx, (size_t)(Lk_pad * Lq * H * es));
    if (!sbuf || !pbuf) return -1;

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
            ctx->stats.ext[0] += (uint64_t)n;
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
            ctx->stats.ext[0] += (uint64_t)n;
        }
        launches += n;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// supports_op
// ------------------------------------------------------------------------------------------------
// test knob (option "debug_refuse_op"): report one ggml op as unsupported, so that the host's own ggml_backend_sched fallback
// (src/core/ggml_extend.hpp:2198-2225: B200 -> CPU -> B200 splits with cross-backend copies) can be exercised against this backend
static std::atomic<int> g_refuse_op{-1};

static void b200_debug_refuse_op(int op) { g_refuse_op.store(op, std::memory_order_relaxed); }

bool b200_supports_op(const b200_device_info&, const ggml_tensor* op) {
    if ((int)op->op == g_refuse_op.load(std::memory_order_relaxed)) return false;
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
        case GGML_OP_IM2COL_3D:
            return s1->type == GGML_TYPE_F32 && s1->nb[0] == 4 && (op->type == GGML_TYPE_F16 || op->type == GGML_TYPE_F32) && ggml_is_contiguous(op);
        case GGML_OP_MUL_MAT: {
            if (!is_f32(op) || !ggml_is_contiguous(op)) return false;
            // Q8_0 weights (the Wan config): dequantised once per weight tensor into a cached f16 copy, then the f16 tensor-core path
            if (s0->type == GGML_TYPE_Q8_0) return ggml_is_contiguous(s0) && s0->ne[0] % 32 == 0 && is_fp(s1) && s1->ne[2] % s0->ne[2] == 0 && s1->ne[3] % s0->ne[3] == 0;
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
        case GGML_OP_IM2COL_3D:
            return b200_launch_im2col_3d(s, b200_make_td(s1), b200_make_td(t), s0->ne[0], s0->ne[1], s0->ne[2], ggml_get_op_params_i32(t, 9), t->op_params);
        case GGML_OP_MUL_MAT: return op_mul_mat(ctx, t);
        case GGML_OP_FLASH_ATTN_EXT: return op_flash_attn(ctx, t);
        default: return -1;
    }
}

static inline bool node_is_noop(const ggml_tensor* t) {
    return t->op == GGML_OP_NONE || t->op == GGML_OP_RESHAPE || t->op == GGML_OP_VIEW || t->op == GGML_OP_PERMUTE || t->op == GGML_OP_TRANSPOSE ||
           ggml_is_empty(t);
}

// ------------------------------------------------------------------------------------------------
// graph identity: sd.cpp rebuilds the ggml graph on EVERY model call (ggml_extend.hpp:3066-3069) but gallocr hands out
// the same addresses for the same topology, so (ops, shapes, strides, params, addresses) identify a repeat exactly.
// A repeat is replayed as one CUDA graph: ~1.5k kernel launches collapse into one cudaGraphLaunch.
// ------------------------------------------------------------------------------------------------
static inline uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    return h ^ (h >> 32);
}

// The identity of a graph: one record per node (op, type, flags, data address, full ne / nb, op_params) and per source (type, data
// address, full ne / nb).  The 64-bit hash of the record array finds the plan; a hit is confirmed by comparing the arrays word for
// word, so a hash collision or a graph that differs in any of these fields can never replay kernels recorded for another graph.
// The same pass notes nodes that WRITE into a WEIGHTS-usage buffer (the reference's LoRA apply adds into model tensors on the runtime
// backend, lora.hpp:934-937): derived weight copies of those ranges must be dropped (see graph_compute).
struct weight_write { const void* ptr; size_t bytes; };

static uint64_t graph_signature(const ggml_cgraph* g, std::vector<uint64_t>& sig, std::vector<weight_write>* writes) {
    sig.clear();
    sig.reserve((size_t)g->n_nodes * 24);
    sig.push_back((uint64_t)g->n_nodes);
    for (int i = 0; i < g->n_nodes; ++i) {
        const ggml_tensor* t = g->nodes[i];
        sig.push_back(((uint64_t)t->op << 32) | ((uint64_t)t->type << 16) | (uint64_t)(t->flags & 0xffff));
        sig.push_back((uint64_t)(uintptr_t)t->data);
        for (int k = 0; k < 4; ++k) { sig.push_back((uint64_t)t->ne[k]); sig.push_back((uint64_t)t->nb[k]); }
        const uint64_t* op = (const uint64_t*)t->op_params;
        for (size_t k = 0; k < sizeof(t->op_params) / 8; ++k) sig.push_back(op[k]);
        for (int sidx = 0; sidx < GGML_MAX_SRC; ++sidx) {
            const ggml_tensor* sr = t->src[sidx];
            if (!sr) break;
            sig.push_back((uint64_t)(uintptr_t)sr->data ^ ((uint64_t)sr->type << 56));
            for (int k = 0; k < 4; ++k) { sig.push_back((uint64_t)sr->ne[k]); sig.push_back((uint64_t)sr->nb[k]); }
        }
        if (writes && (t->flags & GGML_TENSOR_FLAG_COMPUTE) && !node_is_noop(t)) {
            const ggml_tensor* d = (t->op == GGML_OP_CPY && t->src[1]) ? t->src[1] : t;
            const ggml_tensor* root = d->view_src ? d->view_src : d;
            if (root->buffer && root->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS && d->data) writes->push_back({d->data, ggml_nbytes(d)});
        }
    }
    uint64_t h = 0x1234567ull;
    for (uint64_t w : sig) h = mix64(h, w);
    return h;
}

// ------------------------------------------------------------------------------------------------
// peephole fusion.  The reference's block wrappers emit fixed little chains of nodes around every contraction /
// normalisation (ggml_extend.hpp:1008-1040 Linear, :1131-1171 Conv2d, :1502-1520 GroupNorm, :3897-3945 LayerNorm,
// block.hpp:142 "+ SiLU").  Each chain below is executed as ONE kernel when the intermediate tensors have no other
// consumer in this graph; otherwise (or with option "fusion"=0) every node runs on its own, which is also what the
// single-node graphs of test-backend-ops exercise.
// ------------------------------------------------------------------------------------------------
struct fusion_state {
    std::unordered_map<const ggml_tensor*, int> uses;   // consumer count inside this graph
    std::vector<char> done;                             // node already covered by an earlier fused launch
    // FLASH_ATTN_EXT node index -> (skipped CONT node, strided descriptor the kernel reads Q through instead)
    struct q_bypass { ggml_tensor* cont; b200_td q; int split_b; };     // split_b > 0: q is described as [d, Lq, H, split_b] (batch apart)
    std::unordered_map<int, q_bypass> fa_q;
    // K / V operand (the CPY node's tensor) of a FLASH_ATTN_EXT -> the f16 projection output the GEMM epilogue wrote, described as
    // [d, Lk, H, B] over its [B][Lk][H * d] rows: the permute + CONT + cast chain in between was never executed
    struct kv_direct { b200_td td; const ggml_tensor* cast_dst; };
    std::unordered_map<const ggml_tensor*, kv_direct> fa_kv;
    // projections running on a side stream: `done` must have fired before the main stream executes any node after position `pos`
    // (the allocator's lifetimes are those of in-order execution)
    struct side_job { int pos; cudaEvent_t done; };
    std::vector<side_job> pend;
    // adaLN modulation whose scale / shift vectors are produced BETWEEN the NORM and its MUL in graph order (the first use of a block's
    // Modulation output): MUL node index -> the NORM node that was held back; the fused launch happens at the MUL's turn
    std::unordered_map<int, int> deferred_norm;
};

static void count_uses(const ggml_cgraph* g, fusion_state& fs) {
    fs.uses.reserve((size_t)g->n_nodes * 2);
    for (int i = 0; i < g->n_nodes; ++i)
        for (int s = 0; s < GGML_MAX_SRC; ++s) {
            const ggml_tensor* sr = g->nodes[i]->src[s];
            if (!sr) break;
            fs.uses[sr]++;
        }
    fs.done.assign((size_t)g->n_nodes, 0);
}

static inline bool single_use(const fusion_state& fs, const ggml_tensor* t) {
    if (t->flags & GGML_TENSOR_FLAG_OUTPUT) return false;
    auto it = fs.uses.find(t);
    return it != fs.uses.end() && it->second == 1;
}

static inline bool tensors_overlap(const void* a, size_t na, const void* b, size_t nb) {
    return (const char*)a < (const char*)b + nb && (const char*)b < (const char*)a + na;
}

static inline bool is_f32_vec(const ggml_tensor* t, int64_t n) {
    return t && t->type == GGML_TYPE_F32 && ggml_is_contiguous(t) && ggml_nelements(t) == n;
}

static inline bool is_view_op(const ggml_tensor* t) {
    return t->op == GGML_OP_RESHAPE || t->op == GGML_OP_VIEW || t->op == GGML_OP_PERMUTE || t->op == GGML_OP_TRANSPOSE || t->op == GGML_OP_NONE;
}

// next node index > i that does work (views and covered nodes are skipped); -1 at the end
static inline int next_node(const ggml_cgraph* g, const fusion_state& fs, int i) {
    for (int j = i + 1; j < g->n_nodes; ++j)
        if (!fs.done[j] && !is_view_op(g->nodes[j]) && !ggml_is_empty(g->nodes[j])) return j;
    return -1;
}

// does `v` reach `root` through views that keep root's memory order (same base, contiguous, same element count) and that
// nobody else consumes?  Then a kernel may produce `v`'s value by writing root's flat buffer.
static bool order_preserving_view_of(const fusion_state& fs, const ggml_tensor* v, const ggml_tensor* root) {
    for (int depth = 0; depth < 8; ++depth) {
        if (v == root) return true;
        if (!is_view_op(v) || v->op == GGML_OP_NONE || !v->src[0]) return false;
        if (v->data != root->data || !ggml_is_contiguous(v) || ggml_nelements(v) != ggml_nelements(root)) return false;
        if (!single_use(fs, v->src[0])) return false;
        v = v->src[0];
    }
    return false;
}

static inline bool overlaps_range(const void* a, size_t na, const void* b, size_t nb) {
    return (const char*)a < (const char*)b + nb && (const char*)b < (const char*)a + na;
}

// K / V projection of an attention layer (ggml_ext_attention_ext, ggml_extend.hpp:1340-1400): Linear -> reshape [d, H, L, B] ->
// permute(0,2,1,3) -> CONT -> reshape [d, L, H*B] -> CPY to F16 -> FLASH_ATTN_EXT.  The fused attention kernel reads K / V through any
// 16-byte aligned strides, so the epilogue writes the f16 rows [B][L][H * d] once and the attention reads head h of token l at
// (l * H + h) * d: no f32 projection, no permute copy, no cast pass.  Same values: the CONT copies and the CPY rounds f32 -> f16 (RN),
// exactly what the epilogue's conversion does.  `curv`: tensor holding the projection's value (mm, or its bias ADD); `after`: its node index.
struct kv_match { int jc, jp; const ggml_tensor* cp; const ggml_tensor* cast_dst; int64_t d, L, H, B; };
static bool match_kv_projection(const b200_context* ctx, const ggml_cgraph* g, const fusion_state& fs, const ggml_tensor* mm, const ggml_tensor* curv, int after,
                                kv_match* out) {
    static int kv_enabled = -1;
    if (kv_enabled < 0) { const char* e = getenv("GGML_B200_KV_DIRECT"); kv_enabled = (e && *e) ? atoi(e) : 1; }
    if (!kv_enabled || !ctx->opt_chain_fusion || !ctx->opt_tc_gemm || !ctx->opt_fused_attn) return false;
    if (mm->src[0]->type != GGML_TYPE_F16 || mm->ne[3] != 1 || mm->ne[1] <= 4 || !ggml_is_contiguous(mm) || mm->src[1]->ne[3] != 1) return false;
    const int jc = next_node(g, fs, after);
    const int jp = jc >= 0 ? next_node(g, fs, jc) : -1;
    if (jp < 0 || g->nodes[jc]->op != GGML_OP_CONT || g->nodes[jp]->op != GGML_OP_CPY || (curv->flags & GGML_TENSOR_FLAG_OUTPUT)) return false;
    const ggml_tensor* c = g->nodes[jc];
    const ggml_tensor* cp = g->nodes[jp];
    const ggml_tensor* pv = c->src[0];                      // the permuted view of the projection
    const int64_t Mf = mm->ne[0], L = mm->ne[1], Bn = mm->ne[2];
    if (!(c->type == GGML_TYPE_F32 && ggml_is_contiguous(c) && single_use(fs, c) && (c->flags & GGML_TENSOR_FLAG_COMPUTE) && (cp->flags & GGML_TENSOR_FLAG_COMPUTE) &&
          pv->type == GGML_TYPE_F32 && pv->data == curv->data && ggml_are_same_shape(pv, c)))
        return false;
    // views between the projection and the CONT: one consumer each (nobody else reads the f32 projection)
    {
        const ggml_tensor* v = pv;
        int depth = 0;
        while (v != curv && depth++ < 6) {
            if (!is_view_op(v) || !v->src[0] || !single_use(fs, v) || v->data != curv->data) return false;
            v = v->src[0];
        }
        if (v != curv || !single_use(fs, curv)) return false;
    }
    // geometry: pv = [d, L, H, B] over rows of Mf = H * d features
    const int64_t d = pv->ne[0], H = pv->ne[2];
    if (!(d > 0 && d * H == Mf && pv->ne[1] == L && pv->ne[3] == Bn && pv->nb[0] == 4 && pv->nb[2] == (size_t)d * 4 && pv->nb[1] == (size_t)Mf * 4 &&
          (Bn == 1 || pv->nb[3] == (size_t)Mf * L * 4) && d % 8 == 0 && d <= 192))
        return false;
    const ggml_tensor* cd = cp->src[1];     // the cast's destination
    if (!(cd && cd->type == GGML_TYPE_F16 && ggml_is_contiguous(cd) && cd->ne[0] == d && cd->ne[1] == L && ggml_nelements(cd) == ggml_nelements(c) &&
          order_preserving_view_of(fs, cp->src[0], c) && (cp->src[0] == c || single_use(fs, cp->src[0]))))
        return false;
    {   // one reader of the cast.  ggml_cast makes the node its own src[1] (ggml.c: result->src[1] = result), which counts as a use
        auto it = fs.uses.find(cp);
        const int self = cp->src[1] == cp ? 1 : 0;
        if ((cp->flags & GGML_TENSOR_FLAG_OUTPUT) || it == fs.uses.end() || it->second != 1 + self) return false;
    }
    // its one reader: the K or V operand of an attention node the fused kernel will take (no mask, no ALiBi, d == dv)
    const ggml_tensor* fa = nullptr;
    for (int j = jp + 1; j < g->n_nodes && j < jp + 32; ++j) {
        const ggml_tensor* t = g->nodes[j];
        if (t->op == GGML_OP_FLASH_ATTN_EXT && (t->src[1] == cp || t->src[2] == cp)) { fa = t; break; }
    }
    if (!fa || fa->src[1] == fa->src[2] || fa->src[3]) return false;
    float max_bias;
    memcpy(&max_bias, (const float*)fa->op_params + 1, sizeof(float));
    const ggml_tensor* fq = fa->src[0];
    if (!(max_bias == 0.0f && fq->type == GGML_TYPE_F32 && fq->ne[0] == d && fa->src[1]->ne[0] == d && fa->src[2]->ne[0] == d && fa->src[1]->type == GGML_TYPE_F16 &&
          fa->src[2]->type == GGML_TYPE_F16 && fq->ne[3] == 1 && fq->ne[2] % Bn == 0 && (fq->ne[2] / Bn) % H == 0))
        return false;
    *out = kv_match{jc, jp, cp, cd, d, L, H, Bn};
    return true;
}

static bool same_order_view_of(const ggml_tensor* v, const ggml_tensor* root);
static inline void base_range(const ggml_tensor* t, const char** lo, size_t* n);

// MUL_MAT [-> views] [-> CONT of an order-preserving view] [-> views] [-> ADD bias]
static int try_fuse_mul_mat(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered, const ggml_tensor* src1_pre = nullptr,
                            int pre_act = 0, bool side_only = false, bool allow_geglu = true) {
    ggml_tensor* mm = g->nodes[i];
    if (!ggml_is_contiguous(mm) || (mm->flags & GGML_TENSOR_FLAG_OUTPUT)) return -2;
    std::vector<int> chain;
    const ggml_tensor* cur = mm;    // tensor whose flat buffer holds the running value
    int j = next_node(g, fs, i);
    if (j >= 0) {
        ggml_tensor* c = g->nodes[j];
        if (c->op == GGML_OP_CONT && c->type == GGML_TYPE_F32 && ggml_is_contiguous(c) && (c->flags & GGML_TENSOR_FLAG_COMPUTE) &&
            ggml_nelements(c) == ggml_nelements(mm) && c->src[0] != mm && order_preserving_view_of(fs, c->src[0], mm) && single_use(fs, c->src[0])) {
            chain.push_back(j);
            cur = c;
            j = next_node(g, fs, j);
        }
    }
    mm_fusion fz;
    fz.out = (float*)cur->data;
    fz.src1_pre = src1_pre;
    fz.pre_act = pre_act;
    const int64_t M = mm->ne[0], N = mm->ne[1];
    if (j >= 0) {
        ggml_tensor* add = g->nodes[j];
        if (add->op == GGML_OP_ADD && (add->flags & GGML_TENSOR_FLAG_COMPUTE) && add->type == GGML_TYPE_F32 && ggml_is_contiguous(add) &&
            ggml_nelements(add) == ggml_nelements(mm) && order_preserving_view_of(fs, add->src[0], cur) &&
            (add->data == cur->data || single_use(fs, add->src[0])) && !(cur->flags & GGML_TENSOR_FLAG_OUTPUT)) {
            const ggml_tensor* x = add->src[0];
            const ggml_tensor* bv = add->src[1];
            int mode = 0;
            // Linear bias [M] on [M, ...]; conv bias [1,1,OC,1] on [W,H,OC,1] whose GEMM view is [M = W*H, N = OC]
            if (is_f32_vec(bv, M) && bv->ne[0] == M && x->ne[0] == M) mode = 1;
            else if (is_f32_vec(bv, N) && bv->ne[0] == 1 && bv->ne[1] == 1 && bv->ne[2] == N && x->ne[2] == N && x->ne[0] * x->ne[1] == M && x->ne[3] == 1 &&
                     mm->ne[2] * mm->ne[3] == 1)
                mode = 2;
            if (mode) {
                fz.bias = (const float*)bv->data;
                fz.bias_mode = mode;
                fz.out = (float*)add->data;
                chain.push_back(j);
            }
        }
    }
    // GEGLU (FeedForward of the UNet transformer blocks, block.hpp:182-210): proj [+ bias] -> chunk views x | gate -> CONT(gate) -> GELU ->
    // MUL(x, .) -> net.2.  The projection is bound by its OUTPUT bytes (f32 [2 inner, tokens]: 84 MB at 8192 tokens x 2560, written, read
    // back by the GEGLU pass, 42 MB of products written and packed to 16 bits again).  In GEGLU mode the pair kernel pairs each x feature
    // with its gate feature inside one accumulator tile and its epilogue writes ONLY the 16-bit operand of net.2.
    static int geglu_enabled = -1;
    if (geglu_enabled < 0) { const char* e = getenv("GGML_B200_GEGLU_EPI"); geglu_enabled = (e && *e) ? atoi(e) : 1; }
    const ggml_tensor* geglu_act = nullptr;
    int geglu_done = 0;
    if (geglu_enabled && allow_geglu && !src1_pre && !side_only && ctx->opt_chain_fusion && ctx->opt_tc_gemm && ctx->opt_fusion &&
        (mm->src[0]->type == GGML_TYPE_F16 || mm->src[0]->type == GGML_TYPE_BF16) && mm->ne[1] > 4 && mm->ne[3] == 1 && M % 128 == 0 && fz.bias_mode != 2) {
        const ggml_tensor* curv = chain.empty() ? (const ggml_tensor*)mm : g->nodes[chain.back()];
        const int64_t inner = M / 2;
        const int jc = next_node(g, fs, chain.empty() ? i : chain.back());
        const int jg = jc >= 0 ? next_node(g, fs, jc) : -1;
        const int jm = jg >= 0 ? next_node(g, fs, jg) : -1;
        const int jn = jm >= 0 ? next_node(g, fs, jm) : -1;
        if (jn >= 0 && !(curv->flags & GGML_TENSOR_FLAG_OUTPUT) && ggml_is_contiguous(curv)) {
            const ggml_tensor* c = g->nodes[jc];
            const ggml_tensor* ge = g->nodes[jg];
            const ggml_tensor* mul = g->nodes[jm];
            const ggml_tensor* nm = g->nodes[jn];
            const ggml_tensor* gate = c->op == GGML_OP_CONT ? c->src[0] : nullptr;
            const ggml_tensor* xv = (mul->op == GGML_OP_MUL) ? (mul->src[0] == ge ? mul->src[1] : (mul->src[1] == ge ? mul->src[0] : nullptr)) : nullptr;
            auto half_view = [&](const ggml_tensor* v, int64_t off_elems) {
                // a chunk view [inner, tokens..] of curv's rows, `off_elems` floats into each row, read by one node only
                if (!v || v->op != GGML_OP_VIEW || v->type != GGML_TYPE_F32 || !v->src[0] || !single_use(fs, v)) return false;
                const ggml_tensor* base = v->src[0];
                if (base != curv && !(is_view_op(base) && base->op != GGML_OP_NONE && base->data == curv->data && ggml_is_contiguous(base) &&
                                      ggml_nelements(base) == ggml_nelements(curv) && base->ne[0] == curv->ne[0]))
                    return false;
                return v->data == (const char*)curv->data + off_elems * 4 && v->ne[0] == inner && v->nb[0] == 4 && v->nb[1] == (size_t)M * 4 &&
                       ggml_nelements(v) * 2 == ggml_nelements(curv);
            };
            const bool views_ok = half_view(gate, inner) && half_view(xv, 0) && gate->src[0] == xv->src[0] &&
                                  fs.uses.count(gate->src[0]) && fs.uses.at(gate->src[0]) == 2 && (gate->src[0] == curv || single_use(fs, curv));
            if (views_ok && (c->flags & GGML_TENSOR_FLAG_COMPUTE) && c->type == GGML_TYPE_F32 && ggml_is_contiguous(c) && single_use(fs, c) &&
                ge->op == GGML_OP_UNARY && ggml_get_unary_op(ge) == GGML_UNARY_OP_GELU && ge->src[0] == c && single_use(fs, ge) &&
                (ge->flags & GGML_TENSOR_FLAG_COMPUTE) && ge->type == GGML_TYPE_F32 && ggml_is_contiguous(ge) &&
                (mul->flags & GGML_TENSOR_FLAG_COMPUTE) && mul->type == GGML_TYPE_F32 && ggml_is_contiguous(mul) && single_use(fs, mul) &&
                ggml_are_same_shape(mul, ge) && mul->ne[0] == inner &&
                nm->op == GGML_OP_MUL_MAT && nm->src[1] && (nm->src[0]->type == GGML_TYPE_F16 || nm->src[0]->type == GGML_TYPE_BF16) &&
                nm->src[1]->type == GGML_TYPE_F32 && same_order_view_of(nm->src[1], mul) && (nm->src[1] == mul || single_use(fs, nm->src[1])) &&
                !ctx->pack_cache.count(std::make_pair(nm->src[1], (int)nm->src[0]->type))) {
                void* sh = ws_alloc(ctx, (size_t)ggml_nelements(mul) * 2);
                if (sh) {
                    fz.geglu = inner;
                    fz.d16 = sh; fz.d16_type = (int)nm->src[0]->type; fz.skip_f32 = true; fz.d16_done = &geglu_done;
                    geglu_act = nm->src[1];
                    chain.push_back(jc); chain.push_back(jg); chain.push_back(jm);
                }
            }
        }
    }
    // activation: Linear -> GELU / SiLU (the MLPs of the DiT blocks, flux.hpp Mlp; GELU is the tanh form in ggml): applied by the epilogue
    // on the biased accumulator -- the same f32 expression the unary kernel evaluates -- instead of a separate pass over [features, tokens]
    if (!fz.geglu && !src1_pre && ctx->opt_chain_fusion && ctx->opt_tc_gemm && mm->src[0]->type != GGML_TYPE_F32 && mm->ne[1] > 4) {
        const int ju = chain.empty() ? next_node(g, fs, i) : next_node(g, fs, chain.back());
        const ggml_tensor* curv = chain.empty() ? (const ggml_tensor*)mm : g->nodes[chain.back()];
        if (ju >= 0) {
            ggml_tensor* u = g->nodes[ju];
            if (u->op == GGML_OP_UNARY && (u->flags & GGML_TENSOR_FLAG_COMPUTE) && u->type == GGML_TYPE_F32 && ggml_is_contiguous(u) && u->src[0] == curv &&
                ggml_are_same_shape(u, curv) && (u->data == curv->data || single_use(fs, curv)) && !(curv->flags & GGML_TENSOR_FLAG_OUTPUT)) {
                const ggml_unary_op uo = ggml_get_unary_op(u);
                const int a = uo == GGML_UNARY_OP_GELU ? 2 : (uo == GGML_UNARY_OP_SILU ? 1 : 0);
                if (a) {
                    fz.act = a;
                    fz.out = (float*)u->data;
                    chain.push_back(ju);
                }
            }
        }
    }
    // Linear -> GELU -> Linear (flux.hpp Mlp, mmdit, wan FFN): the activated result's ONLY reader is the next contraction, which wants it
    // rounded to its weight type.  The epilogue writes that 16-bit operand itself and -- nothing else reading the f32 tensor -- skips the
    // f32 store: the MLP-up projection is bound by output bytes (f32 [12288, 4352] = 214 MB per Flux block at ~2.7 TB/s of write bandwidth)
    int d16_done = 0;
    const ggml_tensor* d16_act = nullptr;
    static int d16_enabled = -1;
    if (d16_enabled < 0) { const char* e = getenv("GGML_B200_D16"); d16_enabled = (e && *e) ? atoi(e) : 1; }
    if (d16_enabled && fz.act && !fz.geglu && !chain.empty()) {
        const ggml_tensor* curv = g->nodes[chain.back()];
        const int jm = next_node(g, fs, chain.back());
        if (jm >= 0 && single_use(fs, curv) && ggml_is_contiguous(curv) && mm->ne[2] * mm->ne[3] == 1) {
            const ggml_tensor* nm = g->nodes[jm];
            if (nm->op == GGML_OP_MUL_MAT && nm->src[1] == curv && (nm->src[0]->type == GGML_TYPE_F16 || nm->src[0]->type == GGML_TYPE_BF16) &&
                (curv->ne[0] * 2) % 16 == 0 && !ctx->pack_cache.count(std::make_pair(curv, (int)nm->src[0]->type))) {
                void* sh = ws_alloc(ctx, (size_t)ggml_nelements(curv) * 2);
                if (sh) {
                    fz.d16 = sh; fz.d16_type = (int)nm->src[0]->type; fz.skip_f32 = true; fz.d16_done = &d16_done;
                    d16_act = curv;
                }
            }
        }
    }
    // K / V projection of an attention layer: see match_kv_projection
    int kv_done = 0, kv_cont = -1, kv_cpy = -1;
    const ggml_tensor* kv_cp = nullptr;
    fusion_state::kv_direct kvd;
    if (!fz.act && !fz.d16 && !src1_pre && !fz.geglu) {
        const ggml_tensor* curv = chain.empty() ? (const ggml_tensor*)mm : g->nodes[chain.back()];
        kv_match km;
        if (match_kv_projection(ctx, g, fs, mm, curv, chain.empty() ? i : chain.back(), &km)) {
            void* sh = ws_alloc(ctx, (size_t)ggml_nelements(mm) * 2);
            if (sh) {
                fz.d16 = sh; fz.d16_type = GGML_TYPE_F16; fz.skip_f32 = true; fz.d16_done = &kv_done;
                kv_cont = km.jc; kv_cpy = km.jp; kv_cp = km.cp;
                b200_td td;
                td.data = sh; td.type = GGML_TYPE_F16;
                td.ne[0] = km.d; td.ne[1] = km.L; td.ne[2] = km.H; td.ne[3] = km.B;
                td.nb[0] = 2; td.nb[1] = km.d * km.H * 2; td.nb[2] = km.d * 2; td.nb[3] = km.d * km.H * km.L * 2;
                kvd.td = td; kvd.cast_dst = km.cast_dst;
            }
        }
    }
    if (side_only && !kv_cp) return -2;          // launched ahead of its turn on a side stream: only as an in-place K / V projection
    fz.d16_strict = side_only;
    // gated residual of the DiT blocks: x + gate * Linear(y) (flux.hpp:330-400 DoubleStreamBlock, :470-500 SingleStreamBlock; mmdit, wan):
    // ... -> MUL(value, gate [M,1,1,1]) -> ADD(x, .).  Both steps in the epilogue, each rounded like the node it replaces.
    static int gate_enabled = -1;
    if (gate_enabled < 0) { const char* e = getenv("GGML_B200_GATE_FUSION"); gate_enabled = (e && *e) ? atoi(e) : 1; }
    if (gate_enabled && !src1_pre && !fz.d16 && ctx->opt_chain_fusion && ctx->opt_tc_gemm && mm->src[0]->type != GGML_TYPE_F32 && mm->ne[1] > 4 && mm->ne[3] == 1) {
        const int jg = chain.empty() ? next_node(g, fs, i) : next_node(g, fs, chain.back());
        const ggml_tensor* curv = chain.empty() ? (const ggml_tensor*)mm : g->nodes[chain.back()];
        const int ja = jg >= 0 ? next_node(g, fs, jg) : -1;
        if (ja >= 0 && g->nodes[jg]->op == GGML_OP_MUL && g->nodes[ja]->op == GGML_OP_ADD && !(curv->flags & GGML_TENSOR_FLAG_OUTPUT)) {
            const ggml_tensor* mul = g->nodes[jg];
            const ggml_tensor* add = g->nodes[ja];
            const ggml_tensor* val = nullptr;
            const ggml_tensor* gv = nullptr;
            if (order_preserving_view_of(fs, mul->src[0], curv)) { val = mul->src[0]; gv = mul->src[1]; }
            else if (order_preserving_view_of(fs, mul->src[1], curv)) { val = mul->src[1]; gv = mul->src[0]; }
            const ggml_tensor* r = add->src[0] == mul ? add->src[1] : (add->src[1] == mul ? add->src[0] : nullptr);
            if (val && gv && r && r != mul && (mul->flags & GGML_TENSOR_FLAG_COMPUTE) && (add->flags & GGML_TENSOR_FLAG_COMPUTE) && mul->type == GGML_TYPE_F32 &&
                add->type == GGML_TYPE_F32 && ggml_is_contiguous(mul) && ggml_is_contiguous(add) && ggml_are_same_shape(mul, curv) && ggml_are_same_shape(add, mul) &&
                gv->type == GGML_TYPE_F32 && gv->ne[0] == M && gv->ne[1] * gv->ne[2] * gv->ne[3] == 1 && gv->nb[0] == 4 && !((uintptr_t)gv->data & 3) &&
                (single_use(fs, val) || mul->data == curv->data) && single_use(fs, mul) && !(mul->flags & GGML_TENSOR_FLAG_OUTPUT) &&
                r->type == GGML_TYPE_F32 && ggml_is_contiguous(r) && ggml_are_same_shape(r, add) &&
                // the gate vector is read by every CTA for the whole launch: it must not be where the result goes
                !overlaps_range(add->data, ggml_nbytes(add), gv->data, (size_t)M * 4)) {
                fz.gate = (const float*)gv->data;
                fz.residual = (const float*)r->data;
                fz.out = (float*)add->data;
                chain.push_back(jg);
                chain.push_back(ja);
            }
        }
    }
    // residual: ... -> ADD(value, r) with r a same-shape tensor that already exists (the ADD is the very next work node, so r was
    // produced before this MUL_MAT).  Read in the epilogue of the element it is added to, so in-place adds onto r are fine.
    if (!fz.gate && !fz.geglu) {
        const int jr = chain.empty() ? next_node(g, fs, i) : next_node(g, fs, chain.back());
        const ggml_tensor* curv = chain.empty() ? (const ggml_tensor*)mm : g->nodes[chain.back()];
        if (jr >= 0) {
            ggml_tensor* add = g->nodes[jr];
            if (add->op == GGML_OP_ADD && (add->flags & GGML_TENSOR_FLAG_COMPUTE) && add->type == GGML_TYPE_F32 && ggml_is_contiguous(add) &&
                ggml_nelements(add) == ggml_nelements(mm) && mm->ne[3] == 1 && !(curv->flags & GGML_TENSOR_FLAG_OUTPUT)) {
                const ggml_tensor* r = nullptr;
                const ggml_tensor* val = nullptr;
                if (order_preserving_view_of(fs, add->src[0], curv)) { val = add->src[0]; r = add->src[1]; }
                else if (order_preserving_view_of(fs, add->src[1], curv)) { val = add->src[1]; r = add->src[0]; }
                if (r && r->type == GGML_TYPE_F32 && ggml_is_contiguous(r) && ggml_nelements(r) == ggml_nelements(mm) && ggml_are_same_shape(r, add) &&
                    (single_use(fs, val) || add->data == curv->data)) {
                    fz.residual = (const float*)r->data;
                    fz.out = (float*)add->data;
                    chain.push_back(jr);
                }
            }
        }
    }
    if (chain.empty() && !src1_pre && !kv_cp) return -2;
    if (src1_pre && overlaps_range(fz.out, ggml_nbytes(mm), src1_pre->data, ggml_nbytes(src1_pre))) return -2;
    // the fused kernel writes `out` while other CTAs may still be reading the operands: `out` must not live in memory gallocr
    // recycled from an operand that is dead in graph order (e.g. the im2col matrix) -- run unfused then
    auto overlaps = [](const void* a, size_t na, const void* b, size_t nb) { return (const char*)a < (const char*)b + nb && (const char*)b < (const char*)a + na; };
    // (an f32 activation against 16-bit / Q8_0 weights is read by the GEMM from its PACKED copy in workspace, written by a kernel that runs
    //  before the GEMM on the same stream: the f32 tensor's memory may then be where the result goes -- gallocr does exactly that with the
    //  attention output of a DiT block, which used to throw the whole projection + bias + gate + residual chain back to five kernels.
    //  Few-row problems are the exception: the GEMV reads the f32 rows in place.)
    const bool src1_packed = mm->src[1]->type == GGML_TYPE_F32 && mm->src[0]->type != GGML_TYPE_F32 && mm->ne[1] > 4 && !src1_pre;
    if (overlaps(fz.out, ggml_nbytes(mm), mm->src[0]->data, ggml_nbytes(mm->src[0])) ||
        (!src1_packed && overlaps(fz.out, ggml_nbytes(mm), mm->src[1]->data, ggml_nbytes(mm->src[1]))))
        return -2;
    if (fz.residual) { fz.d16 = nullptr; fz.skip_f32 = false; d16_act = nullptr; kv_cp = nullptr; }     // (a residual after the activation: keep the plain path)
    int n = op_mul_mat(ctx, mm, &fz);
    if (fz.geglu) {
        // the pair kernel declined (nothing launched) or did not confirm the operand: run the chain the plain way
        if (n < 0 || !geglu_done) {
            if (n >= 0) return -1;                        // launched without the operand: cannot be undone -- report
            return try_fuse_mul_mat(ctx, g, fs, i, covered, src1_pre, pre_act, side_only, false);
        }
        const ggml_tensor* a = geglu_act;
        ctx->pack_cache[std::make_pair(a, fz.d16_type)] = operand{fz.d16, fz.d16_type, a->ne[0], a->ne[0] * a->ne[1], a->ne[0] * a->ne[1] * a->ne[2]};
        ctx->stats.ext[4] += 1;        // GEGLU projections whose only output is the next Linear's 16-bit operand
    }
    if (n < 0) return n;
    if (fz.gate) ctx->stats.ext[15] += 1;      // gated residuals applied by a GEMM epilogue
    if (kv_cp && kv_done) {
        fs.fa_kv[kv_cp] = kvd;
        chain.push_back(kv_cont);
        chain.push_back(kv_cpy);
        ctx->stats.ext[13] += 1;       // K / V projections handed to the attention kernel in place
    }
    if (d16_act && d16_done)
        ctx->pack_cache[std::make_pair(d16_act, fz.d16_type)] = operand{fz.d16, fz.d16_type, d16_act->ne[0], d16_act->ne[0] * d16_act->ne[1], d16_act->ne[0] * d16_act->ne[1] * d16_act->ne[2]};
    for (int c : chain) fs.done[c] = 1;
    *covered = (int)chain.size();
    return n;
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM convolution.  The reference lowers Conv2d to IM2COL(F16) -> MUL_MAT -> RESHAPE -> PERMUTE -> CONT -> ADD bias
// (ggml.c:4732-4753, ggml_extend.hpp:1131-1171).  When the shape fits the TMA halo-tile scheme the whole chain runs as
//   [GroupNorm statistics] -> NCHW f32 -> NHWC f16 transform (norm, affine, SiLU, nearest x2 folded in) -> tcgen05 conv GEMM
// and the im2col matrix is never materialised.  Filters are repacked once per weight tensor ([OC][KH][KW][IC]) and cached
// by device address; any host write into the weight buffer drops the cached copy (b200_invalidate_address_range).
// ------------------------------------------------------------------------------------------------
struct packed_weight { void* ptr; size_t src_bytes; int device; int64_t ne[4]; size_t bytes; };
static std::mutex g_pw_mutex;
static std::unordered_map<const void*, packed_weight> g_packed_weights;
static std::atomic<uint64_t> g_pw_bytes{0};
// bumped whenever a derived weight copy is dropped: captured CUDA graphs hold raw pointers to those copies and must not be replayed
// across such an event (a long-lived backend whose model was reloaded at the same addresses)
static std::atomic<uint64_t> g_pw_generation{0};

namespace {
std::atomic<uint64_t> g_bc_us[4];
std::atomic<int64_t> g_bc_last_leave{0};
thread_local int64_t t_bc_enter = 0;
int64_t bc_now() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace
void b200_boundary_clock::enter() {
    const int64_t now = bc_now();
    const int64_t last = g_bc_last_leave.load(std::memory_order_relaxed);
    if (last > 0 && now > last && now - last < 2000000000ll) g_bc_us[3].fetch_add((uint64_t)((now - last) / 1000), std::memory_order_relaxed);
    t_bc_enter = now;
}
void b200_boundary_clock::leave(int category) {
    const int64_t now = bc_now();
    if (t_bc_enter > 0) g_bc_us[category & 3].fetch_add((uint64_t)((now - t_bc_enter) / 1000), std::memory_order_relaxed);
    g_bc_last_leave.store(now, std::memory_order_relaxed);
}
void b200_boundary_clock::cut() { g_bc_last_leave.store(0, std::memory_order_relaxed); }
uint64_t b200_boundary_clock::us(int what) { return g_bc_us[what & 3].load(std::memory_order_relaxed); }

uint64_t b200_derived_weight_bytes() { return g_pw_bytes.load(std::memory_order_relaxed); }

void b200_invalidate_address_range(int device, const void* ptr, size_t size) {
    std::lock_guard<std::mutex> lock(g_pw_mutex);
    if (g_packed_weights.empty()) return;
    const char* lo = (const char*)ptr;
    const char* hi = lo + size;
    for (auto it = g_packed_weights.begin(); it != g_packed_weights.end();) {
        const char* a = (const char*)it->first;
        if (it->second.device == device && a < hi && a + it->second.src_bytes > lo) {
            cudaFree(it->second.ptr);   // implicit device synchronisation: no kernel can still be reading it
            g_pw_bytes.fetch_sub(it->second.bytes, std::memory_order_relaxed);
            it = g_packed_weights.erase(it);
            g_pw_generation.fetch_add(1, std::memory_order_relaxed);
        } else {
            ++it;
        }
    }
}

// A derived copy may be kept across graph executions only for a CONSTANT of the model: a leaf tensor (no op, not a graph input) that
// lives in a WEIGHTS-usage buffer, possibly seen through views.  A filter computed inside the graph (the reference casts LoRA factors to
// F16 at run time, lora.hpp:724-733) lives in gallocr's compute buffer, whose addresses are recycled: it is repacked on every execution.
static bool is_constant_weight(const ggml_tensor* w) {
    const ggml_tensor* root = w->view_src ? w->view_src : w;
    if (root->op != GGML_OP_NONE || (root->flags & GGML_TENSOR_FLAG_INPUT)) return false;
    return root->buffer && root->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS;
}

static bool pw_matches_dev(const packed_weight& pw, int device, const ggml_tensor* w) {
    return pw.device == device && pw.src_bytes == ggml_nbytes(w) && pw.ne[0] == w->ne[0] && pw.ne[1] == w->ne[1] && pw.ne[2] == w->ne[2] && pw.ne[3] == w->ne[3];
}

// creates (or finds) the persistent derived copy of a constant weight: kind 0 = packed conv filter [OC][KH][KW][IC], kind 1 = Q8_0 -> f16 rows.
// `may_allocate` is false inside a stream capture.  Caller holds no lock.
static const void* derived_weight(int device, cudaStream_t stream, const ggml_tensor* w, int kind, bool may_allocate, int* launches, bool* overflow) {
    std::lock_guard<std::mutex> lock(g_pw_mutex);
    auto it = g_packed_weights.find(w->data);
    if (it != g_packed_weights.end() && pw_matches_dev(it->second, device, w)) return it->second.ptr;
    if (!may_allocate) { if (overflow) *overflow = true; return nullptr; }
    if (it != g_packed_weights.end()) {      // same address, other shape / device: the old copy is stale
        cudaFree(it->second.ptr);
        g_pw_bytes.fetch_sub(it->second.bytes, std::memory_order_relaxed);
        g_packed_weights.erase(it);
        g_pw_generation.fetch_add(1, std::memory_order_relaxed);
    }
    const size_t bytes = kind == 0 ? ggml_nbytes(w) : (size_t)ggml_nelements(w) * 2;
    // byte cap on the derived copies (GGML_B200_DERIVED_CAP_MB, default 96 GB of the 180 GB): past it every copy is dropped -- they are
    // re-derived on demand, and the generation bump keeps captured plans from replaying against freed memory (cudaFree synchronises)
    static int64_t cap = -1;
    if (cap < 0) { const char* e = getenv("GGML_B200_DERIVED_CAP_MB"); cap = ((e && *e) ? (int64_t)atoll(e) : (int64_t)96 * 1024) * 1048576; }
    if ((int64_t)(g_pw_bytes.load(std::memory_order_relaxed) + bytes) > cap && !g_packed_weights.empty()) {
        for (auto& kv : g_packed_weights) cudaFree(kv.second.ptr);
        g_packed_weights.clear();
        g_pw_bytes.store(0, std::memory_order_relaxed);
        g_pw_generation.fetch_add(1, std::memory_order_relaxed);
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    int n = kind == 0 ? b200_launch_pack_conv_weight(stream, w->data, p, (int)w->ne[0], (int)w->ne[1], w->ne[2], w->ne[3])
                      : b200_launch_dequant_q8_0(stream, w->data, p, ggml_nelements(w) / 32);
    if (n < 0) { cudaFree(p); return nullptr; }
    if (launches) *launches += n;
    g_packed_weights[w->data] = packed_weight{p, ggml_nbytes(w), device, {w->ne[0], w->ne[1], w->ne[2], w->ne[3]}, bytes};
    g_pw_bytes.fetch_add(bytes, std::memory_order_relaxed);
    return p;
}

// SURVEY.md 8f-3, weight ingest: called by the buffer's set_tensor for WEIGHTS-usage buffers once a tensor has been uploaded in full.
// Conv filters (F16, 4-D, 3x3) get their K-major [OC][KH][KW][IC] copy and Q8_0 matrices their f16 rows HERE, at load time, instead of
// by a kernel inside the first forward that meets them; 1x1 filters and F16 / BF16 matrices already are the layout the TMA reads.
void b200_ingest_weight(int device, const ggml_tensor* w) {
    if (!w || !w->data || w->view_src) return;
    int kind = -1;
    if (w->type == GGML_TYPE_F16 && ggml_n_dims(w) == 4 && w->ne[0] == w->ne[1] && w->ne[0] == 3 && w->ne[2] % 64 == 0 && ggml_is_contiguous(w)) kind = 0;
    else if (w->type == GGML_TYPE_Q8_0 && ggml_n_dims(w) >= 2 && ggml_is_contiguous(w)) kind = 1;
    if (kind < 0) return;
    derived_weight(device, cudaStreamPerThread, w, kind, true, nullptr, nullptr);
    cudaStreamSynchronize(cudaStreamPerThread);
}

static const void* get_packed_conv_weight(b200_context* ctx, const ggml_tensor* w, int* launches) {
    // a 1x1 filter [1,1,IC,OC] IS the K-major [OC][IC] operand: read in place
    if (w->ne[0] == 1 && w->ne[1] == 1 && ggml_is_contiguous(w)) return w->data;
    if (!is_constant_weight(w)) {
        void* p = ws_alloc(ctx, ggml_nbytes(w));
        if (!p) return nullptr;
        int n = b200_launch_pack_conv_weight(ctx->stream, w->data, p, (int)w->ne[0], (int)w->ne[1], w->ne[2], w->ne[3]);
        if (n < 0) return nullptr;
        *launches += n;
        ctx->stats.ext[3] += 1;
        return p;
    }
    bool overflow = false;
    const void* p = derived_weight(ctx->device, ctx->stream, w, 0, !ctx->capturing, launches, &overflow);
    if (overflow) ctx->capture_overflow = true;
    return p;
}

// Q8_0 weight tensor -> contiguous f16 [K, rows, ...] copy, same cache and invalidation rule as the packed conv filters
static const void* get_dequantised_weight(b200_context* ctx, const ggml_tensor* w, int* launches) {
    const int64_t n = ggml_nelements(w);
    if (!is_constant_weight(w)) {
        void* p = ws_alloc(ctx, (size_t)n * 2);
        if (!p) return nullptr;
        int r = b200_launch_dequant_q8_0(ctx->stream, w->data, p, n / 32);
        if (r < 0) return nullptr;
        *launches += r;
        return p;
    }
    bool overflow = false;
    const void* p = derived_weight(ctx->device, ctx->stream, w, 1, !ctx->capturing, launches, &overflow);
    if (overflow) ctx->capture_overflow = true;
    return p;
}

struct conv_match {
    int i_im2col = -1, i_mm = -1;
    std::vector<int> chain;          // nodes covered besides the starting one
    const ggml_tensor* x = nullptr;  // image [W,H,IC,N] f32
    const ggml_tensor* w = nullptr;  // filter [KW,KH,IC,OC] f16
    float* out = nullptr;
    const float* bias = nullptr;
    const float* residual = nullptr;
    int64_t OW = 0, OH = 0;
    int dil = 1;
    bool tokens_out = false;         // 1x1 conv whose result is only consumed as tokens [OC, OW*OH]: `out` is that CONT's buffer
};

// does node i (IM2COL) start a conv chain the implicit-GEMM kernel can run?  `up` = 2 when the caller feeds the image through
// a folded nearest-x2 upsample (then x is the LOW-resolution tensor and im2col->src[1] is the UPSCALE node)
static bool match_conv(const ggml_cgraph* g, const fusion_state& fs, int i, conv_match* m) {
    const ggml_tensor* im = g->nodes[i];
    if (im->op != GGML_OP_IM2COL || im->type != GGML_TYPE_F16 || !(im->flags & GGML_TENSOR_FLAG_COMPUTE)) return false;
    const int32_t* p = im->op_params;
    if (p[6] != 1) return false;
    const ggml_tensor* w = im->src[0];
    const ggml_tensor* x = im->src[1];
    if (w->type != GGML_TYPE_F16 || !ggml_is_contiguous(w) || x->type != GGML_TYPE_F32 || !ggml_is_contiguous(x)) return false;
    const int64_t IC = x->ne[2], N = x->ne[3], OC = w->ne[3];
    if (w->ne[2] != IC || N < 1) return false;
    const int64_t OW = im->ne[1], OH = im->ne[2];
    if (OW != x->ne[0] || OH != x->ne[1]) return false;
    if (!b200_conv_tc_supported(N, OH, OW, IC, OC, (int)w->ne[1], (int)w->ne[0], p[0], p[1], p[2], p[3], p[4], p[5])) return false;
    if (!single_use(fs, im)) return false;
    int j = next_node(g, fs, i);
    if (j < 0) return false;
    const ggml_tensor* mm = g->nodes[j];
    if (mm->op != GGML_OP_MUL_MAT || !(mm->flags & GGML_TENSOR_FLAG_COMPUTE) || (mm->flags & GGML_TENSOR_FLAG_OUTPUT)) return false;
    if (!order_preserving_view_of(fs, mm->src[0], im)) return false;
    if (mm->src[0] != im && !single_use(fs, mm->src[0])) return false;
    const ggml_tensor* wv = mm->src[1];
    if (wv->data != w->data || wv->type != GGML_TYPE_F16 || !ggml_is_contiguous(wv) || ggml_nelements(wv) != ggml_nelements(w)) return false;
    if (mm->ne[0] != OW * OH * N || mm->ne[1] != OC || mm->ne[2] * mm->ne[3] != 1) return false;
    m->i_im2col = i; m->i_mm = j;
    m->chain.push_back(j);
    m->x = x; m->w = w; m->OW = OW; m->OH = OH; m->dil = p[4];
    const ggml_tensor* cur = mm;
    m->out = (float*)mm->data;
    int k = next_node(g, fs, j);
    if (k >= 0) {
        const ggml_tensor* c = g->nodes[k];
        const bool cont_ok = c->op == GGML_OP_CONT && c->type == GGML_TYPE_F32 && ggml_is_contiguous(c) && (c->flags & GGML_TENSOR_FLAG_COMPUTE) &&
                             ggml_nelements(c) == ggml_nelements(mm) && c->src[0] != mm && single_use(fs, c->src[0]);
        bool layout_ok = false;
        if (cont_ok && N == 1) layout_ok = order_preserving_view_of(fs, c->src[0], mm);
        if (cont_ok && N > 1) {
            // [OW*OH*N, OC] -> reshape [OW, OH, N, OC] -> PERMUTE(0,1,3,2) -> CONT = [OW, OH, OC, N] (ggml.c:4748-4751): exactly the
            // per-image NCHW planes the conv kernel writes, so the permuting copy is absorbed as well
            const ggml_tensor* pv = c->src[0];
            const ggml_tensor* rs = pv->op == GGML_OP_PERMUTE ? pv->src[0] : nullptr;
            layout_ok = rs && order_preserving_view_of(fs, rs, mm) && single_use(fs, rs) && rs->ne[0] == OW && rs->ne[1] == OH && rs->ne[2] == N &&
                        rs->ne[3] == OC && pv->ne[0] == OW && pv->ne[1] == OH && pv->ne[2] == OC && pv->ne[3] == N && pv->nb[0] == rs->nb[0] &&
                        pv->nb[1] == rs->nb[1] && pv->nb[2] == rs->nb[3] && pv->nb[3] == rs->nb[2];
        }
        if (layout_ok) {
            m->chain.push_back(k);
            cur = c;
            m->out = (float*)c->data;
            k = next_node(g, fs, k);
        } else if (N > 1) {
            return false;     // without the permuting CONT the GEMM result is not in the kernel's output layout
        }
    } else if (N > 1) {
        return false;
    }
    if (k >= 0) {
        const ggml_tensor* add = g->nodes[k];
        if (add->op == GGML_OP_ADD && (add->flags & GGML_TENSOR_FLAG_COMPUTE) && add->type == GGML_TYPE_F32 && ggml_is_contiguous(add) &&
            ggml_nelements(add) == ggml_nelements(mm) && order_preserving_view_of(fs, add->src[0], cur) &&
            (add->data == cur->data || single_use(fs, add->src[0])) && !(cur->flags & GGML_TENSOR_FLAG_OUTPUT)) {
            const ggml_tensor* bv = add->src[1];
            if (is_f32_vec(bv, OC) && bv->ne[0] == 1 && bv->ne[1] == 1 && bv->ne[2] == OC && add->src[0]->ne[2] == OC) {
                m->bias = (const float*)bv->data;
                m->out = (float*)add->data;
                m->chain.push_back(k);
                cur = add;
                k = next_node(g, fs, k);
            }
        }
    }
    // 1x1 conv followed by the image -> token transpose (SpatialTransformer proj_in, block.hpp:548-551: PERMUTE(1,2,0,3) + CONT):
    // computed as D[pixel][oc] = W . X^T instead, which IS the token layout -- no transpose kernel
    if (k >= 0 && w->ne[0] == 1 && w->ne[1] == 1 && !(cur->flags & GGML_TENSOR_FLAG_OUTPUT) && single_use(fs, cur)) {
        const ggml_tensor* c = g->nodes[k];
        const ggml_tensor* pv = c->op == GGML_OP_CONT ? c->src[0] : nullptr;
        if (pv && pv->op == GGML_OP_PERMUTE && pv->src[0] == cur && single_use(fs, pv) && c->type == GGML_TYPE_F32 && ggml_is_contiguous(c) &&
            (c->flags & GGML_TENSOR_FLAG_COMPUTE) && ggml_is_contiguous(cur) && cur->ne[0] == OW && cur->ne[1] == OH && cur->ne[2] == OC && cur->ne[3] == N &&
            c->ne[0] == OC && c->ne[1] == OW && c->ne[2] == OH && c->ne[3] == N && pv->nb[0] == cur->nb[2] && pv->nb[1] == cur->nb[0] &&
            pv->nb[2] == cur->nb[1] && pv->nb[3] == cur->nb[3]) {
            m->tokens_out = true;
            m->out = (float*)c->data;
            m->chain.push_back(k);
            return true;
        }
    }
    // residual add (ResBlock skip / transformer residual) read in the conv epilogue
    if (k >= 0) {
        const ggml_tensor* add = g->nodes[k];
        if (add->op == GGML_OP_ADD && (add->flags & GGML_TENSOR_FLAG_COMPUTE) && add->type == GGML_TYPE_F32 && ggml_is_contiguous(add) &&
            ggml_nelements(add) == ggml_nelements(mm) && !(cur->flags & GGML_TENSOR_FLAG_OUTPUT)) {
            const ggml_tensor* r = nullptr;
            const ggml_tensor* val = nullptr;
            if (order_preserving_view_of(fs, add->src[0], cur)) { val = add->src[0]; r = add->src[1]; }
            else if (order_preserving_view_of(fs, add->src[1], cur)) { val = add->src[1]; r = add->src[0]; }
            if (r && r->type == GGML_TYPE_F32 && ggml_is_contiguous(r) && ggml_are_same_shape(r, add) && (single_use(fs, val) || add->data == cur->data)) {
                m->residual = (const float*)r->data;
                m->out = (float*)add->data;
                m->chain.push_back(k);
            }
        }
    }
    return true;
}

struct conv_prologue {
    const ggml_tensor* src = nullptr;   // tensor actually read (f32 NCHW); equals match.x unless a producer chain was folded in
    int up = 1;
    bool norm = false;
    int n_groups = 0;
    float eps = 0.f;
    const float* gw = nullptr;
    const float* gb = nullptr;
    int act = 0;
    const void* ready_nhwc = nullptr;   // the NHWC f16 image already exists (tokens are NHWC): skip the transform
    const float* addv = nullptr;        // [N][C] vector added to src on the fly: the folded `h + emb` broadcast ADD of a ResBlock
};

static int emit_conv(b200_context* ctx, const conv_match& m, const conv_prologue& pro) {
    int launches = 0;
    const ggml_tensor* src = pro.src;
    const int64_t C = m.w->ne[2], N = m.x->ne[3], H = m.OH / pro.up, W = m.OW / pro.up;
    const void* wp = nullptr;
    bool w_fresh = false;
    if (!m.tokens_out) {
        wp = get_packed_conv_weight(ctx, m.w, &launches);
        if (!wp) return -1;
        w_fresh = launches > 0;      // packed by a kernel of this very graph execution: not yet a constant
    }
    const void* shadow = pro.ready_nhwc;
    if (!shadow) {
        void* sh = ws_alloc(ctx, (size_t)(N * m.OH * m.OW * C * 2));
        if (!sh) return -1;
        float* stats = nullptr;
        if (pro.norm) {
            stats = (float*)ws_alloc(ctx, (size_t)(N * pro.n_groups * 2 * sizeof(float)));
            if (!stats) return -1;
            const size_t pb = ctx->gn_counters ? b200_gn_stats_partial_bytes(N, C, H * W, pro.n_groups) : 0;
            void* partial = pb ? ws_alloc(ctx, pb) : nullptr;
            if (pb && !partial) return -1;
            launches += b200_launch_gn_stats(ctx->stream, (const float*)src->data, stats, N, C, H * W, pro.n_groups, pro.eps, partial, ctx->gn_counters, pro.addv);
        }
        int n = b200_launch_to_nhwc_f16(ctx->stream, (const float*)src->data, sh, N, C, H, W, pro.up, stats, pro.n_groups, pro.gw, pro.gb, pro.act, pro.addv);
        if (n < 0) return -1;
        launches += n;
        shadow = sh;
    }
    if (m.tokens_out) {
        // D[pixel][oc] = sum_c W[oc][c] * X[pixel][c]: the 1x1 filter [1,1,IC,OC] is already the K-major [OC][IC] operand
        b200_gemm_args g;
        memset(&g, 0, sizeof(g));
        g.A = m.w->data; g.B = shadow; g.type = GGML_TYPE_F16;
        g.M = m.w->ne[3]; g.N = m.OH * m.OW; g.K = C;
        g.lda = C; g.ldb = C; g.batch = N; g.a_bcast = N;              // one filter matrix for every image of the batch
        g.b_batch_stride = g.N * C; g.d_batch_stride = g.M * g.N;
        g.D = m.out; g.ldd = g.M;
        g.bias = m.bias; g.bias_mode = m.bias ? 1 : 0;
        if (ctx->opt_early_weights && m.w->buffer && m.w->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS) g.early = 1;
        int n = launch_tc(ctx, g);
        if (n < 0) return -1;
        ctx->stats.reserved[4] += 1;
        return launches + n;
    }
    b200_conv_args c;
    memset(&c, 0, sizeof(c));
    c.x_nhwc = shadow; c.w_packed = wp;
    c.N = N; c.H = m.OH; c.W = m.OW; c.C = C; c.OC = m.w->ne[3];
    c.KH = (int)m.w->ne[1]; c.KW = (int)m.w->ne[0];
    c.dil = m.dil; c.pad = c.dil * (c.KH - 1) / 2;
    c.D = m.out; c.bias = m.bias; c.residual = m.residual;
    c.w_const = (ctx->opt_early_weights && !w_fresh) ? 1 : 0;     // at least the NHWC transform precedes this launch in the graph
    c.w_prefetch = (ctx->opt_wprefetch && !w_fresh && !ctx->graph_writes_weights && m.w->buffer && m.w->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS) ? 1 : 0;
    const bool want_peer = ctx->peer.connected && ctx->peer_out && (const void*)m.out == ctx->peer_out && N == 1;
    if (want_peer) {       // compute + collective in one kernel: the epilogue also stores into the peer GPU's mailbox
        c.D2 = (float*)ctx->peer.remote;
        c.d2_seq = ctx->peer.seq(ctx->peer.mailbox);
        c.d2_slot_floats = (int64_t)(ctx->peer.slot_bytes / 4);
    }
    size_t wsb = b200_conv_tc_workspace_bytes(ctx->info, c);
    void* w = wsb ? ws_alloc(ctx, wsb) : nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->opt_kernel_timing) { e0 = kt_event(ctx); e1 = kt_event(ctx); cudaEventRecord(e0, ctx->stream); }
    int n = b200_launch_conv_tc(ctx->stream, ctx->info, c, w, w ? wsb : 0);
    if (n == 2) { n = 1; ctx->stats.ext[5] += 1; if (want_peer) ctx->peer_fused = true; }   // CTA-pair kernel (gemm_tc2.cu)
    if (ctx->opt_kernel_timing) {
        if (n > 0) {
            cudaEventRecord(e1, ctx->stream);
            ctx->kt_pending.push_back({e0, e1, 2.0 * (double)(c.H * c.W) * (double)c.OC * (double)(c.KH * c.KW * c.C) * (double)c.N});
        } else { ctx->kt_free.push_back(e0); ctx->kt_free.push_back(e1); }
    }
    if (n < 0) return -1;
    ctx->stats.tc_gemm_launches += (uint64_t)n;
    ctx->stats.reserved[4] += 1;   // implicit-GEMM convolutions
    return launches + n;
}

// adaLN modulation of the DiT blocks (Flux::modulate flux.hpp:413-428, MMDiT, Wan):  n = NORM(x);  m = MUL(n, scale);  a = ADD(n, m);
// y = ADD(a, shift)  with scale / shift rows of the modulation Linear.  One row-norm launch that also writes the f16 / bf16 operand of
// the projection that consumes y.
// j1_given >= 0: the MUL's index when the NORM was held back (see try_defer_modulate); dry_run: check the pattern only, launch nothing
static int try_fuse_modulate(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered, int j1_given = -1, bool dry_run = false) {
    ggml_tensor* nrm = g->nodes[i];
    if (nrm->op != GGML_OP_NORM || nrm->type != GGML_TYPE_F32 || !ggml_is_contiguous(nrm) || nrm->src[0]->type != GGML_TYPE_F32 || nrm->src[0]->nb[0] != 4) return -2;
    const int64_t C = nrm->ne[0];
    const int j1 = j1_given >= 0 ? j1_given : next_node(g, fs, i);
    if (j1 < 0) return -2;
    const ggml_tensor* mul = g->nodes[j1];
    if (mul->op != GGML_OP_MUL || !(mul->flags & GGML_TENSOR_FLAG_COMPUTE) || mul->src[0] != nrm || !is_f32_vec(mul->src[1], C) || mul->src[1]->ne[0] != C ||
        !ggml_are_same_shape(mul, nrm) || !single_use(fs, mul))
        return -2;
    const int j2 = next_node(g, fs, j1);
    if (j2 < 0) return -2;
    const ggml_tensor* add1 = g->nodes[j2];
    if (add1->op != GGML_OP_ADD || !(add1->flags & GGML_TENSOR_FLAG_COMPUTE) || add1->src[0] != nrm || add1->src[1] != mul || !single_use(fs, add1) ||
        !ggml_are_same_shape(add1, nrm))
        return -2;
    const int j3 = next_node(g, fs, j2);
    if (j3 < 0) return -2;
    ggml_tensor* add2 = g->nodes[j3];
    if (add2->op != GGML_OP_ADD || !(add2->flags & GGML_TENSOR_FLAG_COMPUTE) || add2->src[0] != add1 || !is_f32_vec(add2->src[1], C) || add2->src[1]->ne[0] != C ||
        !ggml_are_same_shape(add2, nrm) || !ggml_is_contiguous(add2) || add2->type != GGML_TYPE_F32)
        return -2;
    // nrm is read by MUL and ADD only
    auto it = fs.uses.find(nrm);
    if (it == fs.uses.end() || it->second != 2 || (nrm->flags & GGML_TENSOR_FLAG_OUTPUT)) return -2;
    if (add2->data != nrm->src[0]->data && tensors_overlap(add2->data, ggml_nbytes(add2), nrm->src[0]->data, ggml_nbytes(nrm->src[0]))) return -2;
    // the modulation vectors are produced earlier in the graph: their memory must not be where we write
    if (tensors_overlap(add2->data, ggml_nbytes(add2), mul->src[1]->data, (size_t)C * 4) || tensors_overlap(add2->data, ggml_nbytes(add2), add2->src[1]->data, (size_t)C * 4))
        return -2;
    if (dry_run) return 0;
    float eps;
    memcpy(&eps, nrm->op_params, 4);
    int want = -1;
    if (C % 8 == 0) {
        for (int u = j3 + 1; u < g->n_nodes && u < j3 + 64; ++u) {
            const ggml_tensor* c = g->nodes[u];
            if (c->op == GGML_OP_MUL_MAT && c->src[1] == add2 && (c->src[0]->type == GGML_TYPE_F16 || c->src[0]->type == GGML_TYPE_BF16)) {
                want = (int)c->src[0]->type;
                break;
            }
        }
    }
    void* shadow = want >= 0 ? ws_alloc(ctx, (size_t)(ggml_nelements(add2) * 2)) : nullptr;
    int n = b200_launch_norm(ctx->stream, B200_NORM_LAYER, b200_make_td(nrm->src[0]), b200_make_td(add2), eps, (const float*)mul->src[1]->data,
                             (const float*)add2->src[1]->data, shadow, want, 1);
    if (n < 0) return -2;
    if (shadow) ctx->pack_cache[std::make_pair((const ggml_tensor*)add2, want)] = operand{shadow, want, add2->ne[0], add2->ne[0] * add2->ne[1], add2->ne[0] * add2->ne[1] * add2->ne[2]};
    if (j1_given < 0) fs.done[j1] = 1;          // (held-back form: j1 is the node being executed)
    fs.done[j2] = 1; fs.done[j3] = 1;
    *covered = j1_given < 0 ? 3 : 2;
    return n;
}

// NORM whose first reader comes a few nodes later: in the reference's graphs the scale / shift vectors of a block's FIRST modulate are
// emitted between the NORM and the MUL (depth-first node order: ADD(x, MUL(x, scale)) visits the norm, then the Modulation Linear).
// The NORM is held back -- nothing is launched at its turn -- and the fused norm + modulate kernel runs at the MUL's turn, when the
// vectors exist.  Safe because the kernel then reads the NORM's INPUT: no node in between may write into that memory (gallocr may
// have recycled it for one of them once the NORM, its last reader in graph order, has run), and none may read the NORM's output.
static int try_defer_modulate(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i) {
    static int en = -1;
    if (en < 0) { const char* e = getenv("GGML_B200_DEFER_MODULATE"); en = (e && *e) ? atoi(e) : 1; }
    if (!en) return -2;
    ggml_tensor* nrm = g->nodes[i];
    if (nrm->op != GGML_OP_NORM || (nrm->flags & GGML_TENSOR_FLAG_OUTPUT)) return -2;
    const ggml_tensor* x = nrm->src[0];
    const char* xlo; size_t xn;
    base_range(x, &xlo, &xn);
    int j1 = -1, work = 0;
    for (int j = i + 1; j < g->n_nodes && j < i + 48; ++j) {
        const ggml_tensor* t = g->nodes[j];
        bool reads = false;
        for (int sidx = 0; sidx < GGML_MAX_SRC && t->src[sidx]; ++sidx) reads |= t->src[sidx] == nrm;
        if (reads) { j1 = j; break; }
        if (fs.done[j] || is_view_op(t) || ggml_is_empty(t)) continue;
        if (++work > 12) return -2;
        if (tensors_overlap(t->data, ggml_nbytes(t), xlo, xn)) return -2;          // would overwrite the input before the held-back read
        if (t->op == GGML_OP_CPY && t->src[1] && tensors_overlap(t->src[1]->data, ggml_nbytes(t->src[1]), xlo, xn)) return -2;
    }
    if (j1 < 0 || work == 0 || g->nodes[j1]->op != GGML_OP_MUL || is_view_op(g->nodes[j1])) return -2;
    int cov = 0;
    if (try_fuse_modulate(ctx, g, fs, i, &cov, j1, true) != 0) return -2;          // the MUL / ADD / ADD pattern must hold from j1 on
    fs.deferred_norm[j1] = i;
    return 0;                                                                      // nothing launched now
}

// GROUP_NORM -> MUL w -> ADD b [-> SILU]     /     NORM -> MUL w -> ADD b
static int try_fuse_norm(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered, const ggml_tensor* pre_add = nullptr) {
    ggml_tensor* nrm = g->nodes[i];
    const bool group = nrm->op == GGML_OP_GROUP_NORM;
    if (nrm->type != GGML_TYPE_F32 || !ggml_is_contiguous(nrm) || !ggml_is_contiguous(nrm->src[0])) return -2;
    const int64_t nvec = group ? nrm->ne[2] : nrm->ne[0];
    int j = next_node(g, fs, i);
    if (j < 0) return -2;
    ggml_tensor* mul = g->nodes[j];
    if (mul->op != GGML_OP_MUL || mul->src[0] != nrm || !(mul->data == nrm->data || single_use(fs, nrm)) || !is_f32_vec(mul->src[1], nvec)) return -2;
    if (group ? !(mul->src[1]->ne[2] == nvec) : !(mul->src[1]->ne[0] == nvec)) return -2;
    if (!ggml_is_contiguous(mul) || mul->type != GGML_TYPE_F32) return -2;
    int k = next_node(g, fs, j);
    if (k < 0) return -2;
    ggml_tensor* add = g->nodes[k];
    if (add->op != GGML_OP_ADD || add->src[0] != mul || !(add->data == mul->data || single_use(fs, mul)) || !is_f32_vec(add->src[1], nvec)) return -2;
    if (group ? !(add->src[1]->ne[2] == nvec) : !(add->src[1]->ne[0] == nvec)) return -2;
    if (!ggml_is_contiguous(add) || add->type != GGML_TYPE_F32) return -2;
    if ((nrm->flags | mul->flags) & GGML_TENSOR_FLAG_OUTPUT) return -2;
    ggml_tensor* last = add;
    int act = 0, ncov = 2;
    std::vector<int> chain = {j, k};
    if (group) {
        int u = next_node(g, fs, k);
        if (u >= 0 && g->nodes[u]->op == GGML_OP_UNARY && ggml_get_unary_op(g->nodes[u]) == GGML_UNARY_OP_SILU && g->nodes[u]->src[0] == add &&
            (g->nodes[u]->data == add->data || single_use(fs, add)) && !(add->flags & GGML_TENSOR_FLAG_OUTPUT) && ggml_is_contiguous(g->nodes[u]) &&
            g->nodes[u]->type == GGML_TYPE_F32) {
            last = g->nodes[u];
            act = 1;
            chain.push_back(u);
            ncov = 3;
        }
    }
    int n;
    if (group && ctx->opt_tc_gemm && ctx->opt_implicit_conv && single_use(fs, last) && !(last->flags & GGML_TENSOR_FLAG_OUTPUT)) {
        // ... -> IM2COL -> MUL_MAT ...: the normalised activation is only ever read by a convolution: never materialise it
        int ic = next_node(g, fs, chain.back());
        conv_match cm;
        if (ic >= 0 && g->nodes[ic]->op == GGML_OP_IM2COL && g->nodes[ic]->src[1] == last && match_conv(g, fs, ic, &cm)) {
            conv_prologue pro;
            pro.src = pre_add ? pre_add->src[0] : nrm->src[0];
            pro.addv = pre_add ? (const float*)pre_add->src[1]->data : nullptr;
            pro.norm = true;
            pro.n_groups = ggml_get_op_params_i32(nrm, 0);
            memcpy(&pro.eps, (const float*)nrm->op_params + 1, 4);
            pro.gw = (const float*)mul->src[1]->data;
            pro.gb = (const float*)add->src[1]->data;
            pro.act = act;
            n = emit_conv(ctx, cm, pro);
            if (n >= 0) {
                for (int c : chain) fs.done[c] = 1;
                fs.done[ic] = 1;
                for (int c : cm.chain) fs.done[c] = 1;
                *covered = ncov + 1 + (int)cm.chain.size();
                return n;
            }
            if (ctx->capture_overflow) return -1;
        }
    }
    if (pre_add) return -2;            // only the conv prologue can fold the broadcast ADD: the caller runs the ADD and comes back
    // the fused kernel writes `last` while reading the norm's input: identical placement is fine (each CTA owns its group / row),
    // a partial overlap (recycled memory at another offset) is not
    if (last->data != nrm->src[0]->data && tensors_overlap(last->data, ggml_nbytes(last), nrm->src[0]->data, ggml_nbytes(nrm->src[0]))) return -2;
    b200_td dst = b200_make_td(last);
    if (group) {
        float eps;
        memcpy(&eps, (const float*)nrm->op_params + 1, 4);
        n = b200_launch_group_norm(ctx->stream, b200_make_td(nrm->src[0]), dst, ggml_get_op_params_i32(nrm, 0), eps, (const float*)mul->src[1]->data,
                                   (const float*)add->src[1]->data, act);
    } else {
        float eps;
        memcpy(&eps, nrm->op_params, 4);
        // consumers that are F16/BF16-weight contractions want this activation rounded to their type (oracle: ggml-cpu.c:1430-1513):
        // write that copy from the same kernel and hand it to prepare_operand through the pack cache
        int want = -1;
        if (last->ne[0] % 8 == 0 && ggml_is_contiguous(last)) {
            for (int u = i + 1; u < g->n_nodes && u < i + 64; ++u) {
                const ggml_tensor* c = g->nodes[u];
                if (c->op == GGML_OP_MUL_MAT && c->src[1] == last && (c->src[0]->type == GGML_TYPE_F16 || c->src[0]->type == GGML_TYPE_BF16)) {
                    want = (int)c->src[0]->type;
                    break;
                }
            }
        }
        void* shadow = nullptr;
        if (want >= 0) shadow = ws_alloc(ctx, (size_t)(ggml_nelements(last) * 2));
        n = b200_launch_norm(ctx->stream, B200_NORM_LAYER, b200_make_td(nrm->src[0]), dst, eps, (const float*)mul->src[1]->data,
                             (const float*)add->src[1]->data, shadow, want);
        if (n >= 0 && shadow) ctx->pack_cache[std::make_pair((const ggml_tensor*)last, want)] = operand{shadow, want, last->ne[0], last->ne[0] * last->ne[1], last->ne[0] * last->ne[1] * last->ne[2]};
    }
    if (n < 0) return n;
    for (int c : chain) fs.done[c] = 1;
    *covered = ncov;
    return n;
}

// CONT (strided f32 view) -> [order-preserving views] -> CPY to contiguous F16/BF16: one strided converting pass
// (the K / V operands of ggml_ext_attention_ext: cont(permute(x)) followed by ggml_cast(F16), ggml_extend.hpp:1374-1406)
static int try_fuse_cont_cast(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    ggml_tensor* c = g->nodes[i];
    const ggml_tensor* src = c->src[0];
    if (c->type != GGML_TYPE_F32 || src->type != GGML_TYPE_F32 || !ggml_is_contiguous(c) || (c->flags & GGML_TENSOR_FLAG_OUTPUT)) return -2;
    if (src->nb[0] != 4 || c->ne[0] % 8 != 0 || !single_use(fs, c)) return -2;
    if (src->ne[0] != c->ne[0] || src->ne[1] != c->ne[1] || src->ne[2] != c->ne[2] || src->ne[3] != c->ne[3]) return -2;
    int j = next_node(g, fs, i);
    if (j < 0) return -2;
    ggml_tensor* cp = g->nodes[j];
    if (cp->op != GGML_OP_CPY || !(cp->flags & GGML_TENSOR_FLAG_COMPUTE)) return -2;
    const ggml_tensor* dst = cp->src[1];
    if (!(dst->type == GGML_TYPE_F16 || dst->type == GGML_TYPE_BF16) || !ggml_is_contiguous(dst) || ggml_nelements(dst) != ggml_nelements(c)) return -2;
    if (dst->ne[0] != c->ne[0]) return -2;   // rows must stay rows: the cast may only regroup the outer dims
    if (!order_preserving_view_of(fs, cp->src[0], c)) return -2;
    if (cp->src[0] != c && !single_use(fs, cp->src[0])) return -2;
    if (((uintptr_t)src->data & 15) || (src->nb[1] & 15) || (src->nb[2] & 15) || (src->nb[3] & 15) || ((uintptr_t)dst->data & 15)) return -2;
    // gallocr may have placed the cast's destination in the memory of the CONT's source (dead after the CONT in graph order):
    // a single pass that reads one while writing the other would race
    if (tensors_overlap(dst->data, ggml_nbytes(dst), src->data, ggml_nbytes(src))) return -2;
    int n = b200_launch_pack_rows(ctx->stream, b200_make_td(src), dst->data, (int)dst->type, c->ne[0]);
    if (n < 0) return -2;
    fs.done[j] = 1;
    *covered = 1;
    return n;
}

// `v` reaches `root` through views that keep root's flat element order (read-only alias: no consumer-count conditions)
static bool same_order_view_of(const ggml_tensor* v, const ggml_tensor* root) {
    for (int depth = 0; depth < 8; ++depth) {
        if (v == root) return true;
        if (!is_view_op(v) || v->op == GGML_OP_NONE || !v->src[0]) return false;
        if (v->data != root->data || !ggml_is_contiguous(v) || ggml_nelements(v) != ggml_nelements(root)) return false;
        v = v->src[0];
    }
    return false;
}

// memory a (possibly strided) view can touch: the whole tensor it is a view of
static inline void base_range(const ggml_tensor* t, const char** lo, size_t* n) {
    const ggml_tensor* b = t->view_src ? t->view_src : t;
    *lo = (const char*)b->data;
    *n = ggml_nbytes(b);
}

// If the next working node after j is a MUL_MAT with F16 weights whose activation operand is `produced` (f32, contiguous; possibly
// through order-keeping views), return that operand: its producer can write the f16 copy itself and register it in the pack cache.
static const ggml_tensor* next_mm_activation(b200_context* ctx, const ggml_cgraph* g, const fusion_state& fs, int j, const ggml_tensor* produced) {
    if (!ctx->opt_tc_gemm || !ggml_is_contiguous(produced) || produced->type != GGML_TYPE_F32) return nullptr;
    const int nn = next_node(g, fs, j);
    if (nn < 0) return nullptr;
    const ggml_tensor* mm = g->nodes[nn];
    if (mm->op != GGML_OP_MUL_MAT || !mm->src[0] || mm->src[0]->type != GGML_TYPE_F16) return nullptr;
    const ggml_tensor* act = mm->src[1];
    if (!act || act->type != GGML_TYPE_F32 || !same_order_view_of(act, produced)) return nullptr;
    if ((act->ne[0] * 2) % 16) return nullptr;
    if (ctx->pack_cache.count(std::make_pair(act, (int)GGML_TYPE_F16))) return nullptr;
    return act;
}

// CONT(gate half) -> GELU (in place) -> MUL(x half, .) [-> f16 operand of the next Linear]: the GEGLU tail of FeedForward
// (src/model/common/block.hpp:194-207) as one pass over the projection output
static int try_fuse_geglu(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    ggml_tensor* c = g->nodes[i];
    const ggml_tensor* gate = c->src[0];
    if (c->type != GGML_TYPE_F32 || gate->type != GGML_TYPE_F32 || !ggml_is_contiguous(c) || !single_use(fs, c)) return -2;
    if (!ggml_are_same_shape(c, gate)) return -2;
    const int j1 = next_node(g, fs, i);
    if (j1 < 0) return -2;
    ggml_tensor* ge = g->nodes[j1];
    if (ge->op != GGML_OP_UNARY || ggml_get_unary_op(ge) != GGML_UNARY_OP_GELU || ge->src[0] != c || !single_use(fs, ge)) return -2;
    if (!(ge->flags & GGML_TENSOR_FLAG_COMPUTE) || ge->type != GGML_TYPE_F32 || !ggml_is_contiguous(ge)) return -2;
    const int j2 = next_node(g, fs, j1);
    if (j2 < 0) return -2;
    ggml_tensor* mul = g->nodes[j2];
    if (mul->op != GGML_OP_MUL || !(mul->flags & GGML_TENSOR_FLAG_COMPUTE) || mul->type != GGML_TYPE_F32 || !ggml_is_contiguous(mul)) return -2;
    const ggml_tensor* x = mul->src[0] == ge ? mul->src[1] : (mul->src[1] == ge ? mul->src[0] : nullptr);
    if (!x || x == ge || x->type != GGML_TYPE_F32 || !ggml_are_same_shape(x, mul) || !ggml_are_same_shape(ge, mul)) return -2;
    // one pass reads x and gate while it writes mul: the destination must not have been placed over either source
    const char *lo; size_t nb;
    base_range(x, &lo, &nb);
    if (tensors_overlap(mul->data, ggml_nbytes(mul), lo, nb)) return -2;
    base_range(gate, &lo, &nb);
    if (tensors_overlap(mul->data, ggml_nbytes(mul), lo, nb)) return -2;
    const ggml_tensor* act = next_mm_activation(ctx, g, fs, j2, mul);
    void* shadow = act ? ws_alloc(ctx, (size_t)ggml_nelements(mul) * 2) : nullptr;
    int n = b200_launch_geglu(ctx->stream, b200_make_td(x), b200_make_td(gate), b200_make_td(mul), shadow);
    if (n < 0 && shadow) { shadow = nullptr; n = b200_launch_geglu(ctx->stream, b200_make_td(x), b200_make_td(gate), b200_make_td(mul), nullptr); }
    if (n < 0) return -2;
    if (shadow) ctx->pack_cache[std::make_pair(act, (int)GGML_TYPE_F16)] = operand{shadow, GGML_TYPE_F16, act->ne[0], act->ne[0] * act->ne[1], act->ne[0] * act->ne[1] * act->ne[2]};
    fs.done[j1] = 1;
    fs.done[j2] = 1;
    *covered = 2;
    return n;
}

// CONT(permute(q)) whose only consumer is the Q operand of a FLASH_ATTN_EXT (ggml_ext_attention_ext, ggml_extend.hpp:1374-1380): the
// fused attention kernel reads Q rows through any stride, so the copy is skipped and the permuted view handed over instead --
// provided nothing between here and the attention node (nor its output) was placed over the view's memory by the allocator.
static int try_skip_q_cont(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i) {
    if (!ctx->opt_tc_gemm || !ctx->opt_fused_attn) return -2;
    ggml_tensor* c = g->nodes[i];
    const ggml_tensor* src = c->src[0];
    if (c->type != GGML_TYPE_F32 || src->type != GGML_TYPE_F32 || !ggml_is_contiguous(c) || !ggml_are_same_shape(c, src)) return -2;
    if (src->nb[0] != 4 || ((uintptr_t)src->data & 3) || c->ne[0] % 8 || c->ne[0] > 192) return -2;
    int jf = -1;
    for (int j = i + 1; j < g->n_nodes && j < i + 32; ++j) {
        const ggml_tensor* t = g->nodes[j];
        if (t->op == GGML_OP_FLASH_ATTN_EXT && !fs.done[j] && same_order_view_of(t->src[0], c)) { jf = j; break; }
    }
    if (jf < 0) return -2;
    ggml_tensor* fa = g->nodes[jf];
    const ggml_tensor* q = fa->src[0];
    if (!order_preserving_view_of(fs, q, c)) return -2;     // c and the views up to q have exactly one consumer each
    if (q != c && !single_use(fs, q)) return -2;
    if (fa->src[1]->type != GGML_TYPE_F16 || fa->src[2]->type != GGML_TYPE_F16 || fa->src[1]->ne[0] != fa->src[2]->ne[0]) return -2;
    float max_bias;
    memcpy(&max_bias, (const float*)fa->op_params + 1, sizeof(float));
    if (max_bias != 0.0f) return -2;
    // q is [d, Lq, H(, N)] over c's flat order: express it over src's strides
    int split_b = 0;
    b200_td qt = b200_make_td(q);
    if (q->ne[0] != c->ne[0] || q->ne[1] != c->ne[1]) return -2;
    if (q->ne[2] == c->ne[2] && q->ne[3] == c->ne[3]) {
        qt.nb[1] = src->nb[1]; qt.nb[2] = src->nb[2]; qt.nb[3] = src->nb[3];
    } else if (q->ne[3] == 1 && q->ne[2] == c->ne[2] * c->ne[3] && (c->ne[3] == 1 || src->nb[3] == src->nb[2] * (size_t)c->ne[2])) {
        qt.nb[1] = src->nb[1]; qt.nb[2] = src->nb[2]; qt.nb[3] = src->nb[2] * (size_t)q->ne[2];
    } else if (q->ne[3] == 1 && q->ne[2] == c->ne[2] * c->ne[3] && c->ne[3] > 1 && !fa->src[3]) {
        // heads interleaved inside a token row AND a batch (the CFG-batched UNet: [d, H, L, B] projection rows): no single stride walks
        // the merged head-and-batch index, so the view is handed over with the batch as its own dimension
        qt.ne[2] = c->ne[2]; qt.ne[3] = c->ne[3];
        qt.nb[1] = src->nb[1]; qt.nb[2] = src->nb[2]; qt.nb[3] = src->nb[3];
        split_b = (int)c->ne[3];
    } else {
        return -2;
    }
    qt.data = src->data;
    const char* lo; size_t nb;
    base_range(src, &lo, &nb);
    for (int j = i + 1; j <= jf; ++j) {
        const ggml_tensor* t = g->nodes[j];
        if (is_view_op(t) || ggml_is_empty(t)) continue;
        // nodes ahead of this one that are already done are the layer's in-place K / V projections (launched on the side streams beside
        // the Q projection) and their CONT / CPY chains: they wrote workspace only, the allocator's placement of their tensors is moot
        if (fs.done[j]) continue;
        if (tensors_overlap(t->data, ggml_nbytes(t), lo, nb)) return -2;
        if (t->op == GGML_OP_CPY && t->src[1] && tensors_overlap(t->src[1]->data, ggml_nbytes(t->src[1]), lo, nb)) return -2;
    }
    fs.fa_q[jf] = fusion_state::q_bypass{c, qt, split_b};
    return 0;
}

static int try_fuse_flash_attn(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    ggml_tensor* fa = g->nodes[i];
    fa_fusion fz;
    auto it = fs.fa_q.find(i);
    if (it != fs.fa_q.end()) { fz.q_td = &it->second.q; fz.q_cont = it->second.cont; fz.q_split_b = it->second.split_b; }
    auto ik = fs.fa_kv.find(fa->src[1]);
    if (ik != fs.fa_kv.end()) fz.k_td = &ik->second.td;
    auto iv = fs.fa_kv.find(fa->src[2]);
    if (iv != fs.fa_kv.end()) fz.v_td = &iv->second.td;
    if (!(fa->flags & GGML_TENSOR_FLAG_OUTPUT)) fz.shadow_act = next_mm_activation(ctx, g, fs, i, fa);
    static int out_enabled = -1;
    if (out_enabled < 0) { const char* e = getenv("GGML_B200_FA_OUT16"); out_enabled = (e && *e) ? atoi(e) : 1; }
    if (out_enabled && fz.shadow_act && single_use(fs, fa)) {
        // to_out reads the result directly (through views): when nobody else does and the projection takes the packed f16 rows
        // (tensor-core path: more than 4 activation rows or batched), the f32 result is never stored
        const ggml_tensor* act = fz.shadow_act;
        const int jm = next_node(g, fs, i);
        const ggml_tensor* mm = jm >= 0 ? g->nodes[jm] : nullptr;
        if (mm && mm->op == GGML_OP_MUL_MAT && mm->src[1] == act && mm->src[0]->type == GGML_TYPE_F16 && ctx->opt_tc_gemm && ctx->opt_fused_attn &&
            (act->ne[1] > 4 || act->ne[2] * act->ne[3] > 1) && (act == fa || (single_use(fs, act) && order_preserving_view_of(fs, act, fa))))
            fz.out_skip = true;
    }
    // FLASH_ATTN_EXT [d, H*B, Lq] -> VIEW [d, H, Lq, B] -> CONT -> reshape [H*d, Lq, B] -> Linear (to_out, ggml_extend.hpp:1400-1416): when that
    // projection is the only reader and takes its activation as f16 rows anyway, the attention kernel writes exactly those rows and
    // neither the f32 result nor the CONT's copy of it exists
    int jcont = -1;
    if (out_enabled && !fz.shadow_act && !(fa->flags & GGML_TENSOR_FLAG_OUTPUT) && single_use(fs, fa) && ctx->opt_tc_gemm && ctx->opt_fused_attn) {
        const int jc = next_node(g, fs, i);
        if (jc >= 0 && g->nodes[jc]->op == GGML_OP_CONT) {
            const ggml_tensor* c = g->nodes[jc];
            const ggml_tensor* vw = c->src[0];
            const int64_t dv = fa->ne[0], HB = fa->ne[1], Lq = fa->ne[2];
            bool ok = (c->flags & GGML_TENSOR_FLAG_COMPUTE) && c->type == GGML_TYPE_F32 && ggml_is_contiguous(c) && !(c->flags & GGML_TENSOR_FLAG_OUTPUT) &&
                      fa->ne[3] == 1 && vw != fa && is_view_op(vw) && vw->src[0] == fa && vw->data == fa->data && single_use(fs, vw) && ggml_are_same_shape(vw, c) &&
                      vw->ne[0] == dv && vw->ne[2] == Lq && vw->ne[1] * vw->ne[3] == HB && vw->nb[0] == 4 && vw->nb[1] == fa->nb[1] && vw->nb[2] == fa->nb[2] &&
                      (vw->ne[3] == 1 || vw->nb[3] == fa->nb[1] * (size_t)vw->ne[1]) && (dv * 2) % 16 == 0;
            if (ok) {
                const ggml_tensor* act = next_mm_activation(ctx, g, fs, jc, c);
                const ggml_tensor* mm = act ? g->nodes[next_node(g, fs, jc)] : nullptr;
                // the projection must be one that reads the packed operand: tensor-core path (more than 4 activation rows, or batched),
                // and the CONT feeds nothing else
                if (act && mm && single_use(fs, c) && (act == c || single_use(fs, act)) && order_preserving_view_of(fs, act, c) &&
                    (act->ne[1] > 4 || act->ne[2] * act->ne[3] > 1) && mm->src[0]->type == GGML_TYPE_F16) {
                    fz.out_cont = c; fz.out_act = act; fz.out_h = (int)vw->ne[1]; fz.out_b = (int)vw->ne[3];
                    jcont = jc;
                }
            }
        }
    }
    *covered = 0;
    const int n = op_flash_attn(ctx, fa, &fz);
    if (n >= 0 && fz.out_used && jcont >= 0) { fs.done[jcont] = 1; *covered = 1; }
    return n;
}

// SILU(emb) -> MUL_MAT(W, .) [-> ADD bias]: the per-ResBlock embedding projection (block.hpp:150-156) as one GEMV that applies
// the SiLU while it stages the activation row
static int try_fuse_silu_gemv(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    ggml_tensor* u = g->nodes[i];
    if (!ctx->opt_gemv || ggml_get_unary_op(u) != GGML_UNARY_OP_SILU || u->type != GGML_TYPE_F32 || !single_use(fs, u)) return -2;
    const ggml_tensor* x = u->src[0];
    if (x->type != GGML_TYPE_F32 || !ggml_is_contiguous(x) || !ggml_is_contiguous(u) || !ggml_are_same_shape(x, u)) return -2;
    const int j = next_node(g, fs, i);
    if (j < 0) return -2;
    ggml_tensor* mm = g->nodes[j];
    if (mm->op != GGML_OP_MUL_MAT || mm->src[1] != u || !(mm->flags & GGML_TENSOR_FLAG_COMPUTE)) return -2;
    int cov = 0;
    const int n = try_fuse_mul_mat(ctx, g, fs, j, &cov, x, 1);
    if (n < 0) return -2;
    fs.done[j] = 1;
    *covered = cov + 1;
    return n;
}

// tokens [C, HW] -> PERMUTE(1,0,2,3) -> CONT -> reshape [W,H,C,1] -> 1x1 conv (SpatialTransformer proj_out, block.hpp:565-570): tokens
// are already the NHWC image the implicit-GEMM conv reads, so the transpose and the NCHW->NHWC transform cancel: one f16 pack
static int try_fuse_tokens_conv(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    if (!ctx->opt_tc_gemm || !ctx->opt_implicit_conv) return -2;
    ggml_tensor* c = g->nodes[i];
    const ggml_tensor* pv = c->src[0];
    if (c->type != GGML_TYPE_F32 || !ggml_is_contiguous(c) || pv->op != GGML_OP_PERMUTE || !single_use(fs, c)) return -2;
    const ggml_tensor* t = pv->src[0];
    if (!t || t->type != GGML_TYPE_F32 || !ggml_is_contiguous(t) || t->ne[3] != 1) return -2;
    if (pv->ne[0] != t->ne[1] || pv->ne[1] != t->ne[0] || pv->nb[0] != t->nb[1] || pv->nb[1] != t->nb[0] || pv->data != t->data) return -2;
    if (pv->ne[2] != t->ne[2] || pv->nb[2] != t->nb[2]) return -2;
    const int64_t C = t->ne[0], HW = t->ne[1], NB = t->ne[2];
    const int j = next_node(g, fs, i);
    if (j < 0 || g->nodes[j]->op != GGML_OP_IM2COL) return -2;
    const ggml_tensor* x = g->nodes[j]->src[1];
    if (!x || x->ne[2] != C || x->ne[0] * x->ne[1] != HW || x->ne[3] != NB || !order_preserving_view_of(fs, x, c)) return -2;
    if (x != c && !single_use(fs, x)) return -2;
    conv_match cm;
    if (!match_conv(g, fs, j, &cm) || cm.tokens_out || cm.w->ne[0] != 1 || cm.w->ne[1] != 1 || cm.x != x) return -2;
    if ((C * 2) % 16 || ((uintptr_t)t->data & 15)) return -2;
    void* shadow = ws_alloc(ctx, (size_t)(NB * HW * C * 2));
    if (!shadow) return -1;
    int n0 = b200_launch_pack_rows(ctx->stream, b200_make_td(t), shadow, GGML_TYPE_F16, C);
    if (n0 < 0) return -2;
    conv_prologue pro;
    pro.src = x;
    pro.ready_nhwc = shadow;
    int n = emit_conv(ctx, cm, pro);
    if (n < 0) return -1;     // the pack is already in the stream: do not re-run the chain differently on a half-executed state
    fs.done[j] = 1;
    for (int k : cm.chain) fs.done[k] = 1;
    *covered = 1 + (int)cm.chain.size();
    return n0 + n;
}

// Rope::apply_rope, interleaved variant (src/model/common/rope.hpp:966-1010), as emitted for every q and k of a Flux / Wan block:
//   c1 = CONT(PERMUTE(x, 0,2,1,3))            [d, L, H, N]
//   c2 = CONT(PERMUTE(RESHAPE(c1, [2, d/2, L, H*N]), 3,0,1,2))      even / odd split [d/2, L, H*N, 2]
//   rep_j = REPEAT(RESHAPE(VIEW(c2, half j)))  [2, d/2, L, H*N]     pc = CONT(PERMUTE(pe, 3,0,1,2)); pe_j = VIEW(pc, half j)
//   out = ADD(MUL(rep_0, pe_0), MUL(rep_1, pe_1))  [-> RESHAPE -> CPY F16 for k]
// 8 working nodes, one kernel (kernels/rope.cu).  When x itself is MUL(RMS_NORM(v), w) (QKNorm) the match starts there and folds both.
struct rope_match {
    const ggml_tensor* x = nullptr;      // [d, H, L, N] tensor that is permuted
    const ggml_tensor* pe = nullptr;     // [2, 2, d/2, L]
    ggml_tensor* out = nullptr;          // the ADD node
    std::vector<int> nodes;              // working nodes covered, excluding the starting one
};

static bool match_rope(const ggml_cgraph* g, const fusion_state& fs, int i, rope_match* m) {
    const ggml_tensor* c1 = g->nodes[i];
    if (c1->op != GGML_OP_CONT || c1->type != GGML_TYPE_F32 || !ggml_is_contiguous(c1) || !single_use(fs, c1)) return false;
    const ggml_tensor* p1 = c1->src[0];
    if (p1->op != GGML_OP_PERMUTE || !p1->src[0]) return false;
    const ggml_tensor* x = p1->src[0];
    const int64_t d = x->ne[0], H = x->ne[1], L = x->ne[2], N = x->ne[3];
    if (x->type != GGML_TYPE_F32 || d % 4 || d < 4) return false;
    if (p1->ne[0] != d || p1->ne[1] != L || p1->ne[2] != H || p1->ne[3] != N || p1->nb[0] != x->nb[0] || p1->nb[1] != x->nb[2] || p1->nb[2] != x->nb[1] ||
        p1->nb[3] != x->nb[3] || p1->data != x->data)
        return false;
    auto fl = [&](const ggml_tensor* t) { return (t->flags & GGML_TENSOR_FLAG_COMPUTE) != 0; };
    // c2: even/odd split
    int j = next_node(g, fs, i);
    if (j < 0) return false;
    const ggml_tensor* c2 = g->nodes[j];
    if (c2->op != GGML_OP_CONT || !fl(c2) || !ggml_is_contiguous(c2) || c2->ne[0] != d / 2 || c2->ne[1] != L || c2->ne[2] != H * N || c2->ne[3] != 2) return false;
    const ggml_tensor* p2 = c2->src[0];
    if (p2->op != GGML_OP_PERMUTE || !p2->src[0] || p2->src[0]->op != GGML_OP_RESHAPE || p2->src[0]->src[0] != c1) return false;
    const ggml_tensor* r1 = p2->src[0];
    if (r1->ne[0] != 2 || r1->ne[1] != d / 2 || r1->ne[2] != L || r1->ne[3] != H * N || !ggml_is_contiguous(r1)) return false;
    if (p2->nb[0] != r1->nb[1] || p2->nb[1] != r1->nb[2] || p2->nb[2] != r1->nb[3] || p2->nb[3] != r1->nb[0]) return false;
    m->nodes.push_back(j);
    const size_t half = c2->nb[3];
    auto is_half_repeat = [&](const ggml_tensor* rep, int which) {
        if (rep->op != GGML_OP_REPEAT || !fl(rep) || !ggml_is_contiguous(rep) || rep->ne[0] != 2 || rep->ne[1] != d / 2 || rep->ne[2] != L || rep->ne[3] != H * N) return false;
        const ggml_tensor* rs = rep->src[0];
        if (!rs || rs->op != GGML_OP_RESHAPE || rs->ne[0] != 1 || rs->ne[1] != d / 2 || rs->ne[2] != L || rs->ne[3] != H * N) return false;
        const ggml_tensor* v = rs->src[0];
        if (!v || v->op != GGML_OP_VIEW || v->view_src != c2 || (const char*)v->data != (const char*)c2->data + which * half) return false;
        return v->ne[0] == d / 2 && v->ne[1] == L && v->ne[2] == H * N && v->nb[1] == c2->nb[1] && v->nb[2] == c2->nb[2];
    };
    // rep_0
    j = next_node(g, fs, j);
    if (j < 0 || !is_half_repeat(g->nodes[j], 0) || !single_use(fs, g->nodes[j])) return false;
    const ggml_tensor* rep0 = g->nodes[j];
    m->nodes.push_back(j);
    // pc = CONT(PERMUTE(pe, 3,0,1,2))
    j = next_node(g, fs, j);
    if (j < 0) return false;
    const ggml_tensor* pc = g->nodes[j];
    if (pc->op != GGML_OP_CONT || !fl(pc) || !ggml_is_contiguous(pc) || pc->src[0]->op != GGML_OP_PERMUTE) return false;
    const ggml_tensor* pe = pc->src[0]->src[0];
    const ggml_tensor* pp = pc->src[0];
    if (!pe || pe->type != GGML_TYPE_F32 || !ggml_is_contiguous(pe) || pe->ne[0] != 2 || pe->ne[1] != 2 || pe->ne[2] != d / 2 || pe->ne[3] != L) return false;
    if (pp->ne[0] != 2 || pp->ne[1] != d / 2 || pp->ne[2] != L || pp->ne[3] != 2 || pp->nb[0] != pe->nb[1] || pp->nb[1] != pe->nb[2] || pp->nb[2] != pe->nb[3] ||
        pp->nb[3] != pe->nb[0])
        return false;
    m->nodes.push_back(j);
    const size_t pe_half = pc->nb[3];
    auto is_pe_half = [&](const ggml_tensor* v, int which) {
        return v && v->op == GGML_OP_VIEW && v->view_src == pc && (const char*)v->data == (const char*)pc->data + which * pe_half && v->ne[0] == 2 &&
               v->ne[1] == d / 2 && v->ne[2] == L && v->ne[3] == 1 && v->nb[1] == pc->nb[1] && v->nb[2] == pc->nb[2];
    };
    // m0 = MUL(rep_0, pe_0)
    j = next_node(g, fs, j);
    if (j < 0) return false;
    const ggml_tensor* m0 = g->nodes[j];
    if (m0->op != GGML_OP_MUL || !fl(m0) || m0->src[0] != rep0 || !is_pe_half(m0->src[1], 0) || !ggml_is_contiguous(m0)) return false;
    m->nodes.push_back(j);
    // rep_1, m1
    j = next_node(g, fs, j);
    if (j < 0 || !is_half_repeat(g->nodes[j], 1) || !single_use(fs, g->nodes[j])) return false;
    const ggml_tensor* rep1 = g->nodes[j];
    m->nodes.push_back(j);
    j = next_node(g, fs, j);
    if (j < 0) return false;
    const ggml_tensor* m1 = g->nodes[j];
    if (m1->op != GGML_OP_MUL || !fl(m1) || m1->src[0] != rep1 || !is_pe_half(m1->src[1], 1) || !ggml_is_contiguous(m1)) return false;
    m->nodes.push_back(j);
    j = next_node(g, fs, j);
    if (j < 0) return false;
    ggml_tensor* add = g->nodes[j];
    if (add->op != GGML_OP_ADD || !fl(add) || add->src[0] != m0 || add->src[1] != m1 || !ggml_is_contiguous(add) || add->type != GGML_TYPE_F32) return false;
    if (!single_use(fs, m0) || !single_use(fs, m1)) return false;
    // c2 and pc feed exactly their two half views
    auto uses = [&](const ggml_tensor* t) { auto it = fs.uses.find(t); return it == fs.uses.end() ? 0 : it->second; };
    if (uses(c2) != 2 || uses(pc) != 2 || (c2->flags & GGML_TENSOR_FLAG_OUTPUT) || (pc->flags & GGML_TENSOR_FLAG_OUTPUT)) return false;
    m->nodes.push_back(j);
    m->x = x; m->pe = pe; m->out = add;
    return true;
}

// start: the CONT of the permute (i_c1); rms != nullptr when RMS_NORM + MUL in front of it are folded in (then x = rms->src[0])
static int emit_rope(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, const rope_match& rm, const ggml_tensor* rms, const ggml_tensor* rms_mul,
                     int* covered) {
    const ggml_tensor* x = rms ? rms->src[0] : rm.x;
    const float* w = nullptr;
    float eps = 0.f;
    if (rms) {
        w = (const float*)(rms_mul->src[0] == rms ? rms_mul->src[1] : rms_mul->src[0])->data;
        memcpy(&eps, rms->op_params, sizeof(float));
    }
    // k: ... -> RESHAPE -> CPY to F16 (ggml_ext_attention_ext casts K, ggml_extend.hpp:1388-1392): write the f16 rows directly
    void* out = rm.out->data;
    int out_type = GGML_TYPE_F32;
    int extra = -1;
    const int jn = next_node(g, fs, rm.nodes.back());
    if (jn >= 0 && single_use(fs, rm.out)) {
        const ggml_tensor* cp = g->nodes[jn];
        if (cp->op == GGML_OP_CPY && (cp->flags & GGML_TENSOR_FLAG_COMPUTE) && cp->src[1] && cp->src[1]->type == GGML_TYPE_F16 && ggml_is_contiguous(cp->src[1]) &&
            order_preserving_view_of(fs, cp->src[0], rm.out) && (cp->src[0] == rm.out || single_use(fs, cp->src[0])) && cp->src[1]->ne[0] == x->ne[0] &&
            ggml_nelements(cp->src[1]) == ggml_nelements(rm.out)) {
            out = cp->src[1]->data;
            out_type = GGML_TYPE_F16;
            extra = jn;
        }
    }
    // one pass reads x (strided view of the projection output) while writing out
    const char* lo; size_t nb;
    base_range(x, &lo, &nb);
    const size_t out_bytes = (size_t)ggml_nelements(rm.out) * (out_type == GGML_TYPE_F16 ? 2 : 4);
    if (tensors_overlap(out, out_bytes, lo, nb)) return -2;
    if (tensors_overlap(out, out_bytes, rm.pe->data, ggml_nbytes(rm.pe))) return -2;
    int n = b200_launch_rope(ctx->stream, b200_make_td(x), (const float*)rm.pe->data, out, out_type, w, eps);
    if (n < 0) return -2;
    for (int k : rm.nodes) fs.done[k] = 1;
    if (extra >= 0) fs.done[extra] = 1;
    *covered = (int)rm.nodes.size() + (extra >= 0 ? 1 : 0);
    ctx->stats.reserved[7] += 1;
    return n;
}

static int try_fuse_rope(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    rope_match rm;
    if (!match_rope(g, fs, i, &rm)) return -2;
    return emit_rope(ctx, g, fs, rm, nullptr, nullptr, covered);
}

// RMS_NORM(v) -> MUL(., w[d]) -> [rope chain]: QKNorm + RoPE of a Flux q / k (flux.hpp:213-261,279-295)
static int try_fuse_rms_rope(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    const ggml_tensor* rms = g->nodes[i];
    const ggml_tensor* v = rms->src[0];
    if (rms->type != GGML_TYPE_F32 || v->type != GGML_TYPE_F32 || !single_use(fs, rms) || !ggml_is_contiguous(rms)) return -2;
    const int j1 = next_node(g, fs, i);
    if (j1 < 0) return -2;
    const ggml_tensor* mul = g->nodes[j1];
    if (mul->op != GGML_OP_MUL || !(mul->flags & GGML_TENSOR_FLAG_COMPUTE) || !single_use(fs, mul) || !ggml_is_contiguous(mul)) return -2;
    const ggml_tensor* w = mul->src[0] == rms ? mul->src[1] : (mul->src[1] == rms ? mul->src[0] : nullptr);
    if (!w || w == rms || w->type != GGML_TYPE_F32 || !ggml_is_contiguous(w) || ggml_nelements(w) != rms->ne[0] || w->ne[0] != rms->ne[0]) return -2;
    const int j2 = next_node(g, fs, j1);
    if (j2 < 0) return -2;
    rope_match rm;
    if (!match_rope(g, fs, j2, &rm) || rm.x != mul) return -2;
    int cov = 0;
    const int n = emit_rope(ctx, g, fs, rm, rms, mul, &cov);
    if (n < 0) return -2;
    fs.done[j1] = 1;
    fs.done[j2] = 1;
    *covered = cov + 2;
    return n;
}

// RMS_NORM(v) -> MUL(., w[d]) with nothing to absorb behind it (the QKNorm of a DiT double block, whose q / k are concatenated with the
// other stream before the rotation: flux.hpp SelfAttention::pre_attention): one row-norm launch, product rounded like the MUL node
static int try_fuse_rms_mul(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    static int en = -1;
    if (en < 0) { const char* e = getenv("GGML_B200_RMS_MUL"); en = (e && *e) ? atoi(e) : 1; }
    if (!en) return -2;
    ggml_tensor* rms = g->nodes[i];
    const ggml_tensor* v = rms->src[0];
    if (rms->type != GGML_TYPE_F32 || v->type != GGML_TYPE_F32 || v->nb[0] != 4 || !ggml_is_contiguous(rms) || (rms->flags & GGML_TENSOR_FLAG_OUTPUT)) return -2;
    const int j1 = next_node(g, fs, i);
    if (j1 < 0) return -2;
    ggml_tensor* mul = g->nodes[j1];
    if (mul->op != GGML_OP_MUL || !(mul->flags & GGML_TENSOR_FLAG_COMPUTE) || mul->type != GGML_TYPE_F32 || !ggml_is_contiguous(mul) || !ggml_are_same_shape(mul, rms)) return -2;
    if (mul->src[0] != rms) return -2;                          // value * weight, in the node's own operand order
    const ggml_tensor* w = mul->src[1];
    if (w == rms || w->type != GGML_TYPE_F32 || !ggml_is_contiguous(w) || ggml_nelements(w) != rms->ne[0] || w->ne[0] != rms->ne[0]) return -2;
    if (!(mul->data == rms->data || single_use(fs, rms))) return -2;
    // one pass reads v while it writes mul: identical placement is fine (a row is read completely before it is written), partial overlap is not
    if (mul->data != v->data && tensors_overlap(mul->data, ggml_nbytes(mul), v->data, ggml_nbytes(v))) return -2;
    if (tensors_overlap(mul->data, ggml_nbytes(mul), w->data, ggml_nbytes(w))) return -2;
    float eps;
    memcpy(&eps, rms->op_params, 4);
    const int n = b200_launch_norm(ctx->stream, B200_NORM_RMS, b200_make_td(v), b200_make_td(mul), eps, (const float*)w->data, nullptr, nullptr, -1, 0);
    if (n < 0) return -2;
    fs.done[j1] = 1;
    *covered = 1;
    return n;
}

// IM2COL -> ...   or   UPSCALE(nearest x2) -> IM2COL -> ...
static int try_fuse_conv(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int i, int* covered) {
    ggml_tensor* t = g->nodes[i];
    conv_match cm;
    conv_prologue pro;
    int extra = 0;
    int ic = i;
    if (t->op == GGML_OP_UPSCALE) {
        const int mode = ggml_get_op_params_i32(t, 0);
        if ((mode & 0xFF) != GGML_SCALE_MODE_NEAREST || (mode & ~0xFF)) return -2;
        const ggml_tensor* lo = t->src[0];
        if (lo->type != GGML_TYPE_F32 || !ggml_is_contiguous(lo) || !ggml_is_contiguous(t)) return -2;
        if (t->ne[0] != 2 * lo->ne[0] || t->ne[1] != 2 * lo->ne[1] || t->ne[2] != lo->ne[2] || t->ne[3] != lo->ne[3]) return -2;
        if (!single_use(fs, t)) return -2;
        ic = next_node(g, fs, i);
        if (ic < 0 || g->nodes[ic]->op != GGML_OP_IM2COL || g->nodes[ic]->src[1] != t) return -2;
        pro.src = lo;
        pro.up = 2;
        extra = 1;
    }
    if (!match_conv(g, fs, ic, &cm)) return -2;
    if (!pro.src) pro.src = cm.x;
    int n = emit_conv(ctx, cm, pro);
    if (n < 0) return ctx->capture_overflow ? -1 : -2;
    if (extra) fs.done[ic] = 1;
    for (int c : cm.chain) fs.done[c] = 1;
    *covered = extra + (int)cm.chain.size();
    return n;
}

// ------------------------------------------------------------------------------------------------
// side streams.  An in-place K / V projection (match_kv_projection) writes nothing but workspace, so it may run as soon as its input
// exists -- beside the Q projection of its attention layer, or, when its input is an INPUT of the graph (the text context), at the
// very start of the graph beside conv_in and the first ResBlocks.  These GEMMs are fixed-cost bound (7 us for 0.04 GFLOP: 154 context
// rows) and use a handful of SMs; in line they cost the UNet 64 x 4..7 us per forward.
// ------------------------------------------------------------------------------------------------
static int launch_kv_on_side(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, int j, int which, uint64_t* nodes) {
    cudaStream_t main_stream = ctx->stream, ss = ctx->side[which];
    if (!ss) return -2;
    if (ctx->capturing && ctx->ev_pool.size() - ctx->ev_next < 4) return -2;
    cudaEvent_t ef = side_event(ctx), ed = side_event(ctx);
    if (!ef || !ed) return -2;
    cudaEventRecord(ef, main_stream);
    cudaStreamWaitEvent(ss, ef, 0);
    ctx->stream = ss;
    ctx->on_side = true;
    int cov = 0;
    const int n = try_fuse_mul_mat(ctx, g, fs, j, &cov, nullptr, 0, true);
    ctx->stream = main_stream;
    ctx->on_side = false;
    cudaEventRecord(ed, ss);
    if (n < 0) {                                   // declined: nothing of the node itself was launched; it runs in its turn
        cudaStreamWaitEvent(main_stream, ed, 0);
        return -2;
    }
    fs.done[j] = 1;
    fs.pend.push_back(fusion_state::side_job{j, ed});
    ctx->stats.fused_nodes += (uint64_t)cov;
    ctx->stats.side_launches += 1;
    *nodes += (uint64_t)cov + 1;
    return n;
}

static inline const ggml_tensor* base_tensor(const ggml_tensor* t) {
    while (t->view_src) t = t->view_src;
    return t;
}

// projections of graph inputs: every MUL_MAT whose activation is (a view of) a leaf that no node of this graph writes
static void hoist_input_projections(b200_context* ctx, ggml_cgraph* g, fusion_state& fs, uint64_t* launches, uint64_t* nodes) {
    std::unordered_map<const ggml_tensor*, char> written;
    bool scanned = false;
    for (int j = 0; j < g->n_nodes; ++j) {
        ggml_tensor* mm = g->nodes[j];
        if (mm->op != GGML_OP_MUL_MAT || fs.done[j] || !(mm->flags & GGML_TENSOR_FLAG_COMPUTE) || !mm->src[1]) continue;
        const ggml_tensor* root = base_tensor(mm->src[1]);
        if (root->op != GGML_OP_NONE || !root->data || root->type != GGML_TYPE_F32) continue;
        kv_match km;
        if (!match_kv_projection(ctx, g, fs, mm, mm, j, &km)) continue;
        if (!scanned) {
            for (int k = 0; k < g->n_nodes; ++k) {
                const ggml_tensor* t = g->nodes[k];
                if (t->view_src && !is_view_op(t)) written[base_tensor(t)] = 1;         // CPY destinations, in-place ops
            }
            scanned = true;
        }
        if (written.count(root)) continue;
        const int n = launch_kv_on_side(ctx, g, fs, j, 2, nodes);
        if (n > 0) *launches += (uint64_t)n;
    }
}

static enum ggml_status execute_nodes(b200_context* ctx, ggml_cgraph* cgraph, uint64_t* launches, uint64_t* nodes) {
    fusion_state fs;
    ctx->pack_cache.clear();
    ctx->pack_origins.clear();
    ctx->ev_next = 0;
    ctx->launched_any = false;
    ctx->peer_out = nullptr;
    ctx->peer_fused = false;
    ctx->graph_pushed = false;
    if (ctx->peer.connected) {
        // the tensor to exchange: the graph's result -- its LAST node (what the reference's runner reads back, ggml_extend.hpp:2049,2909) --
        // when it has exactly the mailbox's payload size (the eps prediction)
        if (cgraph->n_nodes > 0) {
            const ggml_tensor* t = cgraph->nodes[cgraph->n_nodes - 1];
            if (t->type == GGML_TYPE_F32 && ggml_is_contiguous(t) && ggml_nbytes(t) == ctx->peer.bytes && t->data) ctx->peer_out = t->data;
        }
    }
    const bool fuse = ctx->opt_fusion && cgraph->n_nodes > 1;
    if (fuse) count_uses(cgraph, fs);
    else fs.done.assign((size_t)cgraph->n_nodes, 0);
    const bool side_on = fuse && ctx->opt_side_streams && ctx->opt_chain_fusion && !ctx->opt_kernel_timing && ctx->side[0] && ctx->side[1] && ctx->side[2];
    if (side_on) hoist_input_projections(ctx, cgraph, fs, launches, nodes);
    for (int i = 0; i < cgraph->n_nodes; ++i) {
        ggml_tensor* t = cgraph->nodes[i];
        if (fs.done[i]) continue;
        if (node_is_noop(t)) continue;
        if ((t->flags & GGML_TENSOR_FLAG_COMPUTE) == 0) continue;
        // side-stream work whose position in the graph lies before this node must be complete: from here on the allocator may have
        // handed its inputs' memory to somebody else, and its consumer (the attention node) comes after it
        for (size_t k = 0; k < fs.pend.size();) {
            if (fs.pend[k].pos < i) { cudaStreamWaitEvent(ctx->stream, fs.pend[k].done, 0); fs.pend[k] = fs.pend.back(); fs.pend.pop_back(); }
            else ++k;
        }
        int n = -2;
        if (fuse) {
            int covered = 0;
            if (t->op == GGML_OP_MUL_MAT && side_on) {
                // Q projection at hand: its sibling K / V projections (same activation operand, in-place candidates) start now on the side streams
                int which = 0;
                bool packed = false;
                for (int j = i + 1; j < cgraph->n_nodes && j <= i + 24 && which < 2; ++j) {
                    ggml_tensor* nj = cgraph->nodes[j];
                    if (nj->op == GGML_OP_FLASH_ATTN_EXT) break;
                    if (fs.done[j] || nj->op != GGML_OP_MUL_MAT || nj->src[1] != t->src[1] || !(nj->flags & GGML_TENSOR_FLAG_COMPUTE)) continue;
                    kv_match km;
                    if (!match_kv_projection(ctx, cgraph, fs, nj, nj, j, &km)) continue;
                    if (!packed) {             // the shared activation operand is packed once, on the main stream, before the fork
                        operand tmp;
                        int l = 0;
                        if (!prepare_operand(ctx, nj->src[1], GGML_TYPE_F16, &tmp, &l)) break;
                        *launches += (uint64_t)l;
                        packed = true;
                    }
                    const int r = launch_kv_on_side(ctx, cgraph, fs, j, which, nodes);
                    if (r >= 0) { *launches += (uint64_t)r; ++which; }
                }
            }
            if (t->op == GGML_OP_MUL_MAT) n = try_fuse_mul_mat(ctx, cgraph, fs, i, &covered);
            else if (t->op == GGML_OP_GROUP_NORM || t->op == GGML_OP_NORM) {
                n = try_fuse_norm(ctx, cgraph, fs, i, &covered);
                if (n == -2 && t->op == GGML_OP_NORM && ctx->opt_chain_fusion) n = try_fuse_modulate(ctx, cgraph, fs, i, &covered);
                if (n == -2 && t->op == GGML_OP_NORM && ctx->opt_chain_fusion) n = try_defer_modulate(ctx, cgraph, fs, i);
            }
            else if (t->op == GGML_OP_MUL && !fs.deferred_norm.empty() && fs.deferred_norm.count(i)) {
                // the held-back NORM of this modulate: fused launch now; if the kernel declines, the NORM runs here (late) and the chain unfused
                const int in = fs.deferred_norm[i];
                fs.deferred_norm.erase(i);
                n = try_fuse_modulate(ctx, cgraph, fs, in, &covered, i);
                if (n < 0) {
                    const int r = run_node(ctx, cgraph->nodes[in]);
                    if (r < 0) n = -1;
                    else { *launches += (uint64_t)r; n = -2; }
                }
            }
            else if ((t->op == GGML_OP_IM2COL || t->op == GGML_OP_UPSCALE) && ctx->opt_tc_gemm && ctx->opt_implicit_conv) n = try_fuse_conv(ctx, cgraph, fs, i, &covered);
            else if (t->op == GGML_OP_CONT) {
                n = try_fuse_cont_cast(ctx, cgraph, fs, i, &covered);
                if (n == -2 && ctx->opt_chain_fusion) n = try_fuse_geglu(ctx, cgraph, fs, i, &covered);
                if (n == -2 && ctx->opt_chain_fusion) n = try_fuse_tokens_conv(ctx, cgraph, fs, i, &covered);
                if (n == -2 && ctx->opt_chain_fusion) n = try_fuse_rope(ctx, cgraph, fs, i, &covered);
                if (n == -2 && ctx->opt_chain_fusion) n = try_skip_q_cont(ctx, cgraph, fs, i);
            } else if (t->op == GGML_OP_RMS_NORM && ctx->opt_chain_fusion) {
                n = try_fuse_rms_rope(ctx, cgraph, fs, i, &covered);
                if (n == -2) n = try_fuse_rms_mul(ctx, cgraph, fs, i, &covered);
            }
            else if (t->op == GGML_OP_UNARY && ctx->opt_chain_fusion) n = try_fuse_silu_gemv(ctx, cgraph, fs, i, &covered);
            else if (t->op == GGML_OP_ADD && ctx->opt_chain_fusion && ctx->opt_tc_gemm && ctx->opt_implicit_conv) {
                // ResBlock: h = conv(...) ; h = ADD(h, emb_out [1,1,C,N]) ; GroupNorm(h) -> SiLU -> conv (block.hpp:142-170).  The broadcast ADD has
                // one consumer, the GroupNorm that the conv prologue absorbs: fold it into the statistics and transform kernels (same f32 add,
                // same rounding), never materialise it
                const ggml_tensor* h0 = t->src[0];
                const ggml_tensor* e0 = t->src[1];
                const int jn = next_node(cgraph, fs, i);
                if (jn >= 0 && cgraph->nodes[jn]->op == GGML_OP_GROUP_NORM && cgraph->nodes[jn]->src[0] == t && single_use(fs, t) && t->type == GGML_TYPE_F32 &&
                    h0->type == GGML_TYPE_F32 && e0->type == GGML_TYPE_F32 && ggml_is_contiguous(h0) && ggml_is_contiguous(e0) && ggml_are_same_shape(t, h0) &&
                    e0->ne[0] == 1 && e0->ne[1] == 1 && e0->ne[2] == h0->ne[2] && e0->ne[3] == h0->ne[3] && (h0->ne[0] * h0->ne[1]) % 4 == 0) {
                    int cov = 0;
                    const int r = try_fuse_norm(ctx, cgraph, fs, jn, &cov, t);
                    if (r >= 0) { fs.done[jn] = 1; covered = cov + 1; n = r; }
                    else if (r == -1) n = -1;
                }
            }
            else if (t->op == GGML_OP_FLASH_ATTN_EXT && ctx->opt_chain_fusion) n = try_fuse_flash_attn(ctx, cgraph, fs, i, &covered);
            if (n >= 0) {
                ctx->stats.fused_nodes += (uint64_t)covered;
                *nodes += (uint64_t)covered;
            }
        }
        if (n == -2) n = run_node(ctx, t);
        if (n < 0) {
            GGML_LOG_ERROR("ggml-b200: node %d (%s, %s) is not executable on this backend\n", i, ggml_op_name(t->op), t->name);
            for (auto& pj : fs.pend) cudaStreamWaitEvent(ctx->stream, pj.done, 0);
            return GGML_STATUS_FAILED;
        }
        *launches += (uint64_t)n;
        *nodes += 1;
        if (n > 0) ctx->launched_any = true;
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
    for (auto& pj : fs.pend) cudaStreamWaitEvent(ctx->stream, pj.done, 0);      // every side branch joins before the graph ends
    fs.pend.clear();
    if (!fs.deferred_norm.empty()) {
        // a held-back NORM whose MUL was never reached (covered by another fusion): its value was never produced -- fail loudly
        GGML_LOG_ERROR("ggml-b200: %zu held-back NORM node(s) were never executed\n", fs.deferred_norm.size());
        return GGML_STATUS_FAILED;
    }
    if (!ctx->capturing && side_on) {       // head-room for the capture of this same graph (no event may be created while capturing)
        while (ctx->ev_pool.size() < ctx->ev_next + 8) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); break; }
            ctx->ev_pool.push_back(e);
        }
    }
    if (ctx->peer_out) {
        auto& pr = ctx->peer;
        if (!ctx->peer_fused) {
            if (b200_launch_peer_push(ctx->stream, ctx->peer_out, pr.remote, pr.seq(pr.mailbox), pr.bytes, pr.slot_bytes) < 0) return GGML_STATUS_FAILED;
            *launches += 1;
        }
        b200_launch_peer_signal_wait(ctx->stream, pr.seq(pr.mailbox), pr.flag(pr.remote), pr.flag(pr.mailbox), pr.err(pr.mailbox), 10.0);
        *launches += 1;
        ctx->graph_pushed = true;
    }
    return GGML_STATUS_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// CFG-split exchange: mailbox life cycle (device code in kernels/peer.cu)
// ------------------------------------------------------------------------------------------------
int b200_peer_create(b200_context* ctx, size_t bytes, void* ipc_handle_out64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!ctx || bytes == 0 || (bytes & 15)) return -1;
    cudaSetDevice(ctx->device);
    b200_peer_close(ctx);
    auto& pr = ctx->peer;
    pr.bytes = bytes;
    pr.slot_bytes = (bytes + 255) & ~(size_t)255;
    const size_t total = 2 * pr.slot_bytes + 512;
    if (cudaMalloc(&pr.mailbox, total) != cudaSuccess) { cudaGetLastError(); pr.mailbox = nullptr; return -1; }
    cudaMemset(pr.mailbox, 0, total);
    cudaDeviceSynchronize();
    if (ipc_handle_out64) {
        cudaIpcMemHandle_t h;
        if (cudaIpcGetMemHandle(&h, pr.mailbox) != cudaSuccess) { cudaGetLastError(); memset(ipc_handle_out64, 0, 64); }
        else memcpy(ipc_handle_out64, &h, 64);
    }
    return 0;
}

int b200_peer_connect(b200_context* ctx, const void* peer_ipc_handle64) {
    if (!ctx || !ctx->peer.mailbox) return -1;
    cudaSetDevice(ctx->device);
    auto& pr = ctx->peer;
    if (!peer_ipc_handle64) {
        pr.remote = pr.mailbox;          // loopback: the rank is its own peer (single-GPU self test of the protocol)
        pr.ipc = false;
    } else {
        cudaIpcMemHandle_t h;
        memcpy(&h, peer_ipc_handle64, 64);
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            fprintf(stderr, "[ggml-b200] cudaIpcOpenMemHandle failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
        pr.remote = (char*)p;
        pr.ipc = true;
    }
    pr.pushes = 0;
    pr.connected = true;
    drop_cuda_graphs(ctx);              // plans captured without the exchange must not be replayed
    ctx->plans.clear();
    return 0;
}

int b200_peer_read(b200_context* ctx, void* host_dst) {
    if (!ctx || !ctx->peer.connected || !host_dst || ctx->peer.pushes == 0) return -1;
    cudaSetDevice(ctx->device);
    auto& pr = ctx->peer;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return -1; }
    unsigned err = 0;
    cudaMemcpy(&err, pr.err(pr.mailbox), 4, cudaMemcpyDeviceToHost);
    if (err) { fprintf(stderr, "[ggml-b200] peer exchange %u timed out: the other rank never arrived\n", err); return -2; }
    const size_t slot = (size_t)(pr.pushes & 1);
    return cudaMemcpy(host_dst, pr.mailbox + slot * pr.slot_bytes, pr.bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

void b200_peer_close(b200_context* ctx) {
    if (!ctx) return;
    auto& pr = ctx->peer;
    if (pr.mailbox || pr.remote) {
        cudaSetDevice(ctx->device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        if (pr.ipc && pr.remote) cudaIpcCloseMemHandle(pr.remote);
        if (pr.mailbox) cudaFree(pr.mailbox);
        if (pr.connected) { drop_cuda_graphs(ctx); ctx->plans.clear(); }
    }
    pr = b200_context::peer_state{};
}

static void drop_cuda_graphs(b200_context* ctx) {
    for (auto& kv : ctx->plans)
        if (kv.second.exec) { cudaGraphExecDestroy(kv.second.exec); kv.second.exec = nullptr; }
}

// counters of kernels that run inside a replayed CUDA graph: the deltas recorded at capture are added at every replay
static void stats_add(b200_stats& dst, const b200_stats& a, const b200_stats& b) {   // dst += a - b
    dst.fused_nodes += a.fused_nodes - b.fused_nodes;
    dst.tc_gemm_launches += a.tc_gemm_launches - b.tc_gemm_launches;
    for (int i = 2; i < 8; ++i) if (i != 3) dst.reserved[i] += a.reserved[i] - b.reserved[i];
    for (int i = 0; i < 16; ++i) if (i != 1 && i != 2 && i != 6) dst.ext[i] += a.ext[i] - b.ext[i];
    dst.side_launches += a.side_launches - b.side_launches;
}

enum ggml_status b200_graph_compute(b200_context* ctx, ggml_cgraph* cgraph) {
    const auto host_t0 = std::chrono::steady_clock::now();
    b200_boundary_clock::enter();
    struct host_timer {
        b200_context* c; std::chrono::steady_clock::time_point t0;
        ~host_timer() {
            c->stats.ext[1] += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
            b200_boundary_clock::leave(2);
        }
    } host_timer_guard{ctx, host_t0};
    B200_CUDA_CHECK(cudaSetDevice(ctx->device));
    // the previous graph's device time is finalised here (the host has synchronised in between: it read the result)
    if ((ctx->opt_timing && ctx->timing_pending) || !ctx->kt_pending.empty()) b200_context_finalize_timing(ctx);
    const size_t ws_chunks_before = ctx->ws.chunks.size();
    ws_begin_graph(ctx);
    if (ctx->ws.chunks.size() != ws_chunks_before) {      // workspace was re-allocated: captured graphs point at freed memory
        drop_cuda_graphs(ctx);
        ctx->ws_generation++;
    }
    const bool timing = ctx->opt_timing && !ctx->timing_pending;
    const bool want_graphs = ctx->opt_cuda_graphs && !ctx->opt_kernel_timing && cgraph->n_nodes >= 16;
    b200_context::plan* pl = nullptr;
    std::vector<weight_write> writes;
    std::vector<uint64_t>& sig = ctx->sig_scratch;
    const uint64_t key = graph_signature(cgraph, sig, &writes);
    ctx->graph_writes_weights = !writes.empty();
    if (!writes.empty()) {
        // this graph modifies model weights in place: derived copies (packed conv filters, dequantised Q8_0) of those ranges are stale
        // after it -- and, for consumers later in this same graph, from the write on.  Drop them now (consumers repack in stream order)
        // and again after the graph has been issued; such a graph is never captured.
        for (auto& w : writes) b200_invalidate_address_range(ctx->device, w.ptr, w.bytes);
        ctx->stats.ext[2] += 1;
    }
    if (want_graphs && writes.empty()) {
        if (ctx->plans.size() > 64) { drop_cuda_graphs(ctx); ctx->plans.clear(); }
        pl = &ctx->plans[key];
        if (pl->seen > 0 && pl->sig != sig) {          // hash collision or a graph differing in a field the hash folded away
            if (pl->exec) { cudaGraphExecDestroy(pl->exec); pl->exec = nullptr; }
            *pl = b200_context::plan{};
        }
        if (pl->seen == 0) pl->sig = sig;
    }
    if (timing) cudaEventRecord(ctx->ev_start, ctx->stream);
    enum ggml_status st = GGML_STATUS_SUCCESS;
    uint64_t launches = 0, nodes = 0;

    if (pl && pl->exec && pl->ws_generation == ctx->ws_generation && pl->pw_generation == g_pw_generation.load(std::memory_order_relaxed)) {
        // ---- replay
        cudaError_t e = cudaGraphLaunch(pl->exec, ctx->stream);
        if (e != cudaSuccess) {
            GGML_LOG_ERROR("ggml-b200: cudaGraphLaunch failed: %s\n", cudaGetErrorString(e));
            return GGML_STATUS_FAILED;
        }
        launches = pl->launches;
        nodes = pl->nodes;
        const b200_stats zero{};
        stats_add(ctx->stats, pl->delta, zero);
        ctx->stats.reserved[3]++;   // CUDA-graph replays
        ctx->graph_pushed = pl->pushed;
    } else if (pl && pl->seen >= 1 && !pl->no_capture && ctx->ws.chunks.size() <= 1) {
        // ---- second sighting (or a stale plan: workspace moved / derived weights dropped): capture while executing nothing, then launch
        //      the instantiated graph
        if (pl->exec) { cudaGraphExecDestroy(pl->exec); pl->exec = nullptr; }
        ctx->capture_overflow = false;
        ctx->capturing = true;
        const b200_stats before = ctx->stats;
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            st = execute_nodes(ctx, cgraph, &launches, &nodes);
            e = cudaStreamEndCapture(ctx->stream, &graph);
        }
        ctx->capturing = false;
        bool ok = e == cudaSuccess && st == GGML_STATUS_SUCCESS && graph && !ctx->capture_overflow;
        if (ok) {
            cudaGraphExec_t exec = nullptr;
            e = cudaGraphInstantiate(&exec, graph, 0);
            ok = e == cudaSuccess && exec;
            if (ok) {
                pl->exec = exec;
                pl->launches = launches;
                pl->nodes = nodes;
                pl->ws_generation = ctx->ws_generation;
                pl->pw_generation = g_pw_generation.load(std::memory_order_relaxed);
                pl->delta = b200_stats{};
                stats_add(pl->delta, ctx->stats, before);
                pl->pushed = ctx->graph_pushed;
                e = cudaGraphLaunch(exec, ctx->stream);
                ok = e == cudaSuccess;
            }
        }
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
            cudaGetLastError();
            ctx->stats = before;                         // nothing of the abandoned capture ran
            pl->no_capture = !ctx->capture_overflow;     // an overflow is cured by the eager run below (it creates the derived weights): capture next time
            if (pl->exec) { cudaGraphExecDestroy(pl->exec); pl->exec = nullptr; }
            if (st != GGML_STATUS_SUCCESS && !ctx->capture_overflow) return st;
            // fall back to eager execution of this call (also when a node needed a one-time allocation that cannot be captured)
            ws_begin_graph(ctx);
            launches = nodes = 0;
            st = execute_nodes(ctx, cgraph, &launches, &nodes);
        }
    } else {
        st = execute_nodes(ctx, cgraph, &launches, &nodes);
        if (pl) pl->seen++;
    }
    if (st != GGML_STATUS_SUCCESS) return st;
    if (!writes.empty()) {
        // copies made DURING this graph from pre-write values (a conv that precedes the write in graph order)
        for (auto& w : writes) b200_invalidate_address_range(ctx->device, w.ptr, w.bytes);
    }
    ctx->stats.ext[6] = g_pw_bytes.load(std::memory_order_relaxed);
    if (ctx->graph_pushed) { ctx->peer.pushes++; ctx->stats.ext[8] += 1; }
    ctx->stats.kernel_launches += launches;
    ctx->stats.nodes_executed += nodes;
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
