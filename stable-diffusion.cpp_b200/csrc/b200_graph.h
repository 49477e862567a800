// b200_graph.h -- per-backend-instance execution context and the graph executor entry points.
#pragma once

#include "b200_common.h"

#include <string>
#include <unordered_map>
#include <utility>
#include <functional>
#include <vector>

struct ggml_cgraph;

// device-memory scratch used inside one graph_compute (operand packing, split-K partials, attention scores).
// Bump allocation; chunks are only released at the start of a later graph, after the stream has drained.
struct b200_workspace {
    struct chunk { char* base; size_t size; size_t used; };
    std::vector<chunk> chunks;
    size_t high_water = 0;   // bytes requested during the current graph
};

// a contraction operand as the tcgen05 GEMM wants it: K-major rows of one element type
struct b200_operand {
    const void* ptr;
    int type;
    int64_t ld;            // row stride, elements
    int64_t batch_stride;  // dim-2 stride, elements
    int64_t b3_stride;     // dim-3 stride, elements
};
struct b200_pack_key_hash {
    size_t operator()(const std::pair<const ggml_tensor*, int>& k) const { return std::hash<const void*>()(k.first) * 31u + (size_t)k.second; }
};

struct b200_context {
    int device = 0;
    b200_device_info info{};
    std::string name;
    cudaStream_t stream = nullptr;
    cudaEvent_t copy_event = nullptr, ev_start = nullptr, ev_stop = nullptr;
    b200_stats stats{};
    b200_workspace ws;
    // options (include/ggml-b200.h: ggml_backend_b200_set_option)
    bool opt_fusion = true;
    bool opt_tc_gemm = true;
    bool opt_timing = true;
    bool opt_cuda_graphs = false;
    bool opt_fused_attn = true;       // single-kernel FLASH_ATTN_EXT (0 = GEMM + softmax + GEMM through workspace)
    bool opt_implicit_conv = true;    // IM2COL+MUL_MAT chains as TMA halo-tile implicit GEMM (0 = materialised im2col)
    bool opt_early_weights = false;   // tcgen05 GEMM fetches its first ring-full of constant weights before the PDL wait
    bool opt_chain_fusion = true;     // GEGLU tail, Q read in place by attention, f16 operand copies written by their producers
    bool opt_gemv = true;             // MUL_MAT with <= 4 activation rows as a weight-streaming GEMV instead of a tcgen05 tile
    bool opt_fold_batch = false;      // MUL_MAT of one weight matrix against a contiguous batch of activations runs as one GEMM with N * batch rows
    bool opt_precise_f32 = true;      // F32 x F32 MUL_MAT as 3xTF32 (hi/lo operand split, three tensor-core passes): f32-class accuracy like the CPU oracle's f32 dot
    bool opt_q8_activations = true;   // Q8_0-weight contractions quantise the activation rows to Q8_0 like the CPU oracle does (q8_0 x q8_0 dot)
    bool opt_kernel_timing = false;   // per-launch CUDA events around every tcgen05 GEMM (roofline pass only)
    bool timing_pending = false;
    struct kt_pair { cudaEvent_t start, stop; double flops; };
    std::vector<kt_pair> kt_pending;
    std::vector<cudaEvent_t> kt_free;
    double kt_us = 0, kt_flops = 0;
    // repeated-graph cache (CUDA graph replay)
    struct plan {
        cudaGraphExec_t exec = nullptr; int seen = 0; bool no_capture = false; uint64_t launches = 0, nodes = 0, ws_generation = 0, pw_generation = 0;
        std::vector<uint64_t> sig;     // full identity of the graph this plan was recorded for (compared word for word on a hash hit)
        b200_stats delta{};            // counters of the kernels inside the captured graph (added at every replay)
        bool pushed = false;           // the captured graph ends with a peer push (kernels/peer.cu)
    };
    std::vector<uint64_t> sig_scratch;
    std::unordered_map<uint64_t, plan> plans;
    uint64_t ws_generation = 0;
    // per-graph-execution cache of packed (type-converted) contraction operands, keyed by ggml tensor node
    std::unordered_map<std::pair<const ggml_tensor*, int>, b200_operand, b200_pack_key_hash> pack_cache;
    // CFG-split exchange over NVLink peer memory (kernels/peer.cu; include/ggml-b200.h ggml_backend_b200_peer_*)
    struct peer_state {
        bool connected = false, ipc = false;
        size_t bytes = 0, slot_bytes = 0;
        char* mailbox = nullptr;        // own: [slot 0 | slot 1 | flag @ 2*slot | seq @ +128 | err @ +256]
        char* remote = nullptr;         // the peer's mailbox mapped here (== mailbox in loopback mode)
        uint64_t pushes = 0;            // pushes issued so far (host count; selects the slot to read)
        unsigned* flag(char* base) const { return (unsigned*)(base + 2 * slot_bytes); }
        unsigned* seq(char* base) const { return (unsigned*)(base + 2 * slot_bytes + 128); }
        unsigned* err(char* base) const { return (unsigned*)(base + 2 * slot_bytes + 256); }
    } peer;
    const void* peer_out = nullptr;     // data pointer of the current graph's output tensor when it is to be pushed to the peer
    bool peer_fused = false;            // ... and its producer has already stored it there (fused epilogue)
    bool graph_pushed = false;          // the graph executed last contained a push
    unsigned* gn_counters = nullptr;   // B200_GN_COUNTERS zeroed counters of the chunked GroupNorm statistics (self-resetting)
    // side streams: in-place K / V projections (workspace-only results) run beside the main stream -- [0], [1] the two projections of the
    // attention layer at hand while the main stream computes Q, [2] projections of graph INPUTS (the text context: 2 x 16 cross-attention
    // layers of the SD1.5 UNet) hoisted to the start of the graph.  Fork / join by events; inside a capture they become graph branches.
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_next = 0;
    bool opt_side_streams = true;
    bool opt_wprefetch = false;        // GEMM / conv CTAs request their constant weight slab from L2 before the PDL wait (b200_gemm_args::wprefetch)
    bool graph_writes_weights = false; // the graph being executed stores into a WEIGHTS buffer: nothing in it is treated as a constant
    bool on_side = false;              // launches currently go to a side stream (ctx->stream is swapped)
    struct pack_origin { cudaStream_t stream; cudaEvent_t ready; };
    std::unordered_map<std::pair<const ggml_tensor*, int>, pack_origin, b200_pack_key_hash> pack_origins;   // packed operands produced on a side stream
    bool capturing = false, capture_overflow = false;
    bool launched_any = false;        // a kernel of the current graph execution has been launched

    ~b200_context();
};

b200_context* b200_context_create(const b200_device_info& info);
int b200_context_set_option(b200_context* ctx, const char* key, int value);
void b200_context_finalize_timing(b200_context* ctx);

// vtable back-ends
enum ggml_status b200_graph_compute(b200_context* ctx, ggml_cgraph* cgraph);
bool b200_supports_op(const b200_device_info& dev, const ggml_tensor* op);

// caches of derived weight layouts are keyed by device address: any host write into a range drops them
void b200_invalidate_address_range(int device, const void* ptr, size_t size);
// weight ingest: derive the layouts the kernels read (packed 3x3 conv filters, Q8_0 -> f16 rows) when a weight is uploaded
void b200_ingest_weight(int device, const ggml_tensor* w);
uint64_t b200_derived_weight_bytes();

// Host-time attribution at the plugin boundary (process-wide, microseconds): every vtable entry that moves data or computes is bracketed,
// and the time BETWEEN two boundary calls is what the host spent outside the backend (the reference's graph rebuild, gallocr, sampler).
struct b200_boundary_clock {
    static void enter();
    static void leave(int category);     // 0 set_tensor, 1 get_tensor (includes waiting for the device), 2 graph_compute (host side)
    static uint64_t us(int what);        // 0..2 as above, 3 = outside the backend
    static void cut();                   // a gap that spans this point is not attributed (called when the counters are read)
};

// CFG-split exchange (kernels/peer.cu)
int b200_peer_create(b200_context* ctx, size_t bytes, void* ipc_handle_out64);
int b200_peer_connect(b200_context* ctx, const void* peer_ipc_handle64);   // nullptr: loopback (single-GPU self test)
int b200_peer_read(b200_context* ctx, void* host_dst);
void b200_peer_close(b200_context* ctx);
