/* ggml-b200.h -- C-ABI of libggml-b200.so, a Blackwell (sm_100a) ggml backend plugin.
 *
 * This is the drop-in boundary: the library is loaded by the reference's own loader
 * (ggml/src/ggml-backend-reg.cpp:221-266, env GGML_BACKEND_PATH or ggml_backend_load(path))
 * and from then on stable-diffusion.cpp's unchanged host code (GGMLRunner, ggml_gallocr,
 * sample(), VAE decode) talks to it only through ggml's vtables
 * (ggml/src/ggml-backend-impl.h:17-230).  It mirrors the public header of the reference's
 * CUDA backend (ggml/include/ggml-cuda.h) entry for entry.
 *
 * Every symbol below is `extern "C"` with plain pointers / integers.  The ggml_* handle types
 * are the reference's own opaque C structs (ggml/include/ggml-backend.h).
 */
#ifndef GGML_B200_H
#define GGML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* opaque handles of the host program (declared in ggml/include/ggml-backend.h) */
struct ggml_backend;
struct ggml_backend_reg;
struct ggml_backend_buffer_type;
typedef struct ggml_backend*             ggml_backend_t;
typedef struct ggml_backend_reg*         ggml_backend_reg_t;
typedef struct ggml_backend_buffer_type* ggml_backend_buffer_type_t;

#define GGML_B200_NAME        "B200"
#define GGML_B200_MAX_DEVICES 16

/* ---- the two symbols the reference's dlopen loader resolves (ggml-backend-reg.cpp:231-266;
 *      macro form GGML_BACKEND_DL_IMPL / GGML_BACKEND_DL_SCORE_IMPL, ggml-backend-impl.h:240-271) */

/* Returns the static registry object (api_version == GGML_BACKEND_API_VERSION == 2). */
ggml_backend_reg_t ggml_backend_init(void);
/* 0 = "cannot run here" (no driver, or no compute-capability-10.0 device); otherwise 100. */
int ggml_backend_score(void);

/* ---- direct entry points (replace ggml_backend_cuda_* of ggml/include/ggml-cuda.h:25-45) */

/* same object ggml_backend_init() returns                       (ggml_backend_cuda_reg, ggml-cuda.h:44) */
ggml_backend_reg_t ggml_backend_b200_reg(void);
/* new backend instance (own stream + workspace) on `device`; NULL on error
 *                                                                (ggml_backend_cuda_init, ggml-cuda.h:25) */
ggml_backend_t ggml_backend_b200_init(int device);
/* true if `backend` was created by this library                 (ggml_backend_is_cuda, ggml-cuda.h:27) */
int ggml_backend_is_b200(ggml_backend_t backend);
/* device-memory buffer type of `device`                         (ggml_backend_cuda_buffer_type, ggml-cuda.h:30) */
ggml_backend_buffer_type_t ggml_backend_b200_buffer_type(int device);
/* pinned host buffer type for staging                           (ggml_backend_cuda_host_buffer_type, ggml-cuda.h:36) */
ggml_backend_buffer_type_t ggml_backend_b200_host_buffer_type(void);
/* number of usable sm_100 devices                               (ggml_backend_cuda_get_device_count, ggml-cuda.h:38) */
int ggml_backend_b200_get_device_count(void);
/* "NVIDIA B200 (sm_100, 148 SMs)" style text                    (ggml_backend_cuda_get_device_description, :39) */
void ggml_backend_b200_get_device_description(int device, char* description, size_t description_size);
/* free / total HBM in bytes                                     (ggml_backend_cuda_get_device_memory, :40) */
void ggml_backend_b200_get_device_memory(int device, size_t* free_bytes, size_t* total_bytes);

/* ---- extensions, also reachable through reg->iface.get_proc_address(reg, "<name>")
 *      (ggml-backend-impl.h:214-224), which is how an unmodified host finds them */

typedef struct ggml_b200_stats {
    uint64_t graphs;            /* graph_compute calls                                       */
    uint64_t kernel_launches;   /* kernels launched (all hand-written: no library dispatch) */
    uint64_t nodes_executed;    /* ggml nodes covered                                        */
    uint64_t fused_nodes;       /* nodes absorbed into a neighbour's kernel                  */
    double   last_graph_ms;     /* device time of the last graph_compute (CUDA events)      */
    double   total_graph_ms;
    uint64_t tc_gemm_launches;  /* tcgen05 GEMM launches                                     */
    uint64_t reserved[8];       /* [0] flop of the tcgen05 GEMMs timed under "kernel_timing", [1] their device time in us,
                                   [2] fused flash-attention launches, [3] CUDA-graph replays, [4] implicit-GEMM convolutions,
                                   [5] attention launches that read Q in place (CONT skipped), [6] few-row GEMV launches, [7] fused RoPE launches */
    uint64_t ext[16];           /* [0] CUDA-core reference GEMM launches (gemm_ref.cu: must stay 0 on every model path; the parity tests assert it),
                                   [1] host microseconds spent inside graph_compute (signature, fusion planning, launches / cudaGraphLaunch),
                                   [2] graphs that wrote into a WEIGHTS buffer (derived weight copies dropped), [3] conv filters packed per graph
                                   (filter computed inside the graph: no persistent copy), [4] GEGLU projections run in the pair kernel's GEGLU mode (only output: the next Linear's 16-bit operand), [5] 2-CTA GEMM launches,
                                   [6] bytes of derived weight copies alive, [7] unfused (GEMM + softmax + GEMM) attention executions,
                                   [8] graphs that ended with a peer exchange (kernels/peer.cu),
                                   [9..12] host microseconds at the plugin boundary, process-wide: inside set_tensor, inside get_tensor (includes
                                   waiting for the device), inside graph_compute (host side), and OUTSIDE the backend between two boundary
                                   calls (the host's own work: graph rebuild, gallocr, sampler),
                                   [13] K / V projections whose f16 rows the attention kernel read in place (permute + CONT + cast never run),
                                   [14] attention launches that wrote only the f16 rows of the output projection (no f32 result, no CONT),
                                   [15] gated residuals (x + gate * Linear(y)) applied by a GEMM epilogue */
    uint64_t side_launches;     /* in-place K / V projections launched on a side stream: beside the Q projection of their layer, or -- for
                                   projections of graph inputs (the text context) -- hoisted to the start of the graph */
} ggml_b200_stats;

/* copy the backend instance's counters; returns 0 on success.  Counters of kernels that run inside a replayed CUDA graph are
 * accounted at every replay from the deltas recorded when the graph was captured, so they stay alive on the measured path. */
int ggml_backend_b200_get_stats(ggml_backend_t backend, ggml_b200_stats* out);
void ggml_backend_b200_reset_stats(ggml_backend_t backend);

/* Run-time options (also read once from the environment, GGML_B200_<KEY>=value):
 *   "fusion"      1/0   graph-level fusion (0 = one kernel per ggml node; used to diff fused vs unfused)
 *   "tc_gemm"     1/0   tcgen05 GEMM (0 = CUDA-core reference GEMM kernel, bring-up/debug only)
 *   "timing"      1/0   record CUDA events around every graph_compute (last_graph_ms)
 *   "cuda_graphs" 1/0   replay captured CUDA graphs for repeated identical ggml graphs
 *   "kernel_timing" 1/0 CUDA events around every tcgen05 GEMM launch (roofline pass; disables graph replay while on)
 *   "fused_attn"  1/0   single-kernel FLASH_ATTN_EXT (0 = GEMM + softmax + GEMM through workspace)
 *   "implicit_conv" 1/0 IM2COL+MUL_MAT chains as TMA halo-tile implicit GEMM (0 = materialised im2col)
 *   "early_weights" 1/0 GEMMs fetch the first tiles of constant weights before the programmatic-dependent-launch wait (default 0)
 *   "chain_fusion" 1/0  producer-side fusions: GEGLU tail, Q read in place, f16 operand copies written by their producers, RoPE, adaLN
 *   "gemv"        1/0   MUL_MAT with <= 4 activation rows as a weight-streaming GEMV
 * returns 0 on success, -1 for an unknown key. */
int ggml_backend_b200_set_option(ggml_backend_t backend, const char* key, int value);

/* Diagnostic, host-only (no reference counterpart; the reference's CUDA backend keeps its kernel selection internal,
 * ggml/src/ggml-cuda/ggml-cuda.cu:1825-2038): the plan -- tile width of the CTA pair, split-K factor, filter taps served per image box --
 * the halo-reuse 3x3 convolution would run for `batch` images of H x W x C -> OC on `sm_count` SMs, and its modelled time in microseconds.
 * Returns 1, or 0 when the shape is outside the halo envelope (W % 8, H % 16, C % 64).  tests/test_cabi.py pins it against the committed
 * hardware sweep. */
int ggml_backend_b200_debug_conv_plan(int64_t batch, int64_t H, int64_t W, int64_t C, int64_t OC, int sm_count, int* bn, int* splits, int* taps,
                                      double* model_us);

/* Diagnostic, host-only: launch geometry of the CTA-pair tcgen05 kernel for a problem of `batch` x [M rows of A, N rows of B, nkb 64-wide k-blocks
 * (halo modes: ring stages)] with tile width bn, split-K factor `splits` and `halo_taps` (0 | 3 | 9) filter taps per image box: ring stages, bytes
 * per stage, dynamic shared memory, TMEM columns, CTAs.  Returns 0 when the plan is outside the kernel's envelope. */
int ggml_backend_b200_debug_pair_geometry(int64_t M, int64_t N, int64_t batch, int nkb, int bn, int splits, int halo_taps, int sm_count, int* stages,
                                          int* stage_bytes, int64_t* smem_bytes, int* tmem_cols, int* ctas);

/* ---- CFG-batch split over a pair of GPUs, one process per GPU (kernels/peer.cu).  The reference offers nothing here
 * (its sample() is serial, stable-diffusion.cpp:2811-2836); the closest reference interface is the meta backend's
 * "ggml_backend_comm_init / _allreduce_tensor" extension pair (ggml/src/ggml-backend-meta.cpp:2207-2220), found the same way:
 * reg->iface.get_proc_address(reg, "ggml_backend_b200_peer_mailbox_*").
 *   create   allocates this backend's mailbox for payloads of `bytes` bytes (the eps prediction) and returns its 64-byte CUDA IPC handle
 *   connect  maps the OTHER rank's mailbox from its handle (NULL = loopback: the rank is its own peer, a single-GPU self test).  From
 *            then on every graph_compute whose output tensor has exactly `bytes` bytes ends with: the tensor stored into the peer's
 *            mailbox over NVLink (by the epilogue of the convolution producing it where that is the CTA-pair kernel, by a copy kernel
 *            otherwise), a sequence number published in the peer's flag, and a device-side wait for the peer's -- all on the backend
 *            stream, inside the captured CUDA graph
 *   read     synchronises the stream and copies the payload the peer stored here to `host_dst`; -2 when the peer never arrived (the
 *            device-side wait gives up after 10 s instead of hanging the GPU)
 * All return 0 on success. */
int  ggml_backend_b200_peer_mailbox_create(ggml_backend_t backend, size_t bytes, void* ipc_handle_out64);
int  ggml_backend_b200_peer_mailbox_connect(ggml_backend_t backend, const void* peer_ipc_handle64);
int  ggml_backend_b200_peer_mailbox_read(ggml_backend_t backend, void* host_dst);
void ggml_backend_b200_peer_mailbox_close(ggml_backend_t backend);

/* ggml_backend_device_i::supports_op (ggml-backend-impl.h:186) without a device: 1 when every B200 of this backend executes `op`.
 * Callable on a machine without a GPU -- the CPU test suite walks the reference's model graphs with it, because any 0 would make
 * sd.cpp route that node to its CPU backend through ggml_backend_sched (src/core/ggml_extend.hpp:2198-2225). */
int ggml_backend_b200_op_supported(const struct ggml_tensor* op);

#ifdef __cplusplus
}
#endif
#endif /* GGML_B200_H */
