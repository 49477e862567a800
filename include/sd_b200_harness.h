/* sd_b200_harness.h -- C-ABI of the synthetic-weight whole-model harness.
 *
 * The harness is HOST code: it instantiates the reference's own, unmodified graph builders
 *   UNetModelRunner          (src/model/diffusion/unet.hpp:747-858)
 *   AutoEncoderKL decode     (src/model/vae/auto_encoder_kl.hpp:589-748)
 *   Flux::FluxRunner         (src/model/diffusion/flux.hpp)
 * and the reference's own sampler (src/runtime/denoiser.hpp:2794 sample_k_diffusion) on any
 * ggml backend in the registry ("CPU" = the oracle, "B200_<i>" = this repo's plugin), with
 * seeded synthetic weights (no checkpoints exist offline).  The same binary drives both
 * backends, so everything above ggml's vtable boundary -- graph construction, gallocr,
 * scheduler/sampler index math -- is bit-identical by construction.
 *
 * Plain pointers and sizes only; used from Python through ctypes (bench.py, tests/).
 * All tensors are float32 in ggml order: ne[0] fastest.  Return codes: 0 = ok, <0 = error
 * (message retrievable through sdh_last_error()).
 */
#ifndef SD_B200_HARNESS_H
#define SD_B200_HARNESS_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdh_model sdh_model;

typedef struct sdh_tensor {
    float*  data;   /* caller-owned; NULL = tensor absent */
    int64_t ne[4];  /* ggml order: ne[0] fastest */
} sdh_tensor;

/* Load a ggml backend plugin (.so exporting ggml_backend_init; ggml-backend-reg.cpp:221-266).
 * Returns the number of devices registered after the load, <0 on failure. */
int sdh_load_backend(const char* so_path);

/* Number of devices in ggml's registry, and their names (e.g. "CPU", "B200_0"). */
int         sdh_device_count(void);
const char* sdh_device_name(int index);

const char* sdh_last_error(void);

/* arch:    "sd15_unet" | "sdxl_unet" | "unet_tiny" | "vae_decoder" | "vae_decoder_sdxl" | "flux_schnell" | "flux_tiny" | "flux_1x1" (one double + one
 *          single block at full FLUX.1 width) |
 *          "mmdit_sd3" (SD3-medium MMDiT) | "wan_1_3b" (Wan2.1-T2V-1.3B DiT) | "clip_l" (CLIP ViT-L/14 text encoder: x = token ids as
 *          floats [n_token, N]) | "wan_vae_decoder" (Wan causal 3-D VAE decoder: x = latent [W, H, T, 16]; with T > 1 the reference's
 *          single-graph decode returns NaN from the second frame on, on its own CPU backend too, so parity is pinned on one frame)
 * wtype:   "f32" | "f16" | "bf16" | "q8_0"  (dtype of Linear weights; conv weights are always F16 as
 *           in the reference, ggml_extend.hpp:3600-3603; norm/bias are F32)
 * flags:   bit0 = flash attention graph variant (--diffusion-fa), bit1 = conv2d direct, bit2 = place the parameters without filling them
 *          (graph / supports_op walks that never run the model: a multi-GB synthetic checkpoint is not generated)
 * seed:    weight seed -- the same seed gives byte-identical weights on every backend
 * n_threads: threads for the CPU backend (ignored by GPU backends)                           */
sdh_model* sdh_model_create(const char* device, const char* arch, const char* wtype,
                            int flags, uint64_t seed, int n_threads);
void       sdh_model_free(sdh_model* m);

/* bytes of parameters resident on the compute backend / number of parameter tensors */
size_t sdh_model_param_bytes(const sdh_model* m);
int    sdh_model_param_count(const sdh_model* m);

/* One forward through GGMLRunner::compute (graph build + gallocr + H2D of inputs + graph_compute
 * + D2H of the result: exactly what sample() pays per model call, ggml_extend.hpp:3152).
 *  unet:  x [W,H,C,N], timesteps [N], context [C_ctx,77,N], y [adm,N] (SDXL) -> out like x
 *  vae:   x = latent [W,H,C,N] -> out [8W,8H,3,N]
 *  flux:  x [W,H,C,N] latent, timesteps [N], context [4096,L_txt,N], y [768,N]
 * out->data must hold the result (query with sdh_model_out_shape first); out->ne is filled.
 * wall_ms (optional) receives the host wall-clock of the call.                                */
int sdh_model_out_shape(sdh_model* m, const sdh_tensor* x, int64_t out_ne[4]);
int sdh_model_forward(sdh_model* m, const sdh_tensor* x, const sdh_tensor* timesteps,
                      const sdh_tensor* context, const sdh_tensor* y, sdh_tensor* out,
                      double* wall_ms);

/* Build the graph for these inputs (no compute) and write one line per node to `path`:
 *   idx op(unary-op) type ne0..3 nb0..3 flags | per src: type ne nb name
 * Returns the node count.  Used to derive supports_op coverage and the algorithmic FLOP count. */
int sdh_model_dump_graph(sdh_model* m, const sdh_tensor* x, const sdh_tensor* timesteps,
                         const sdh_tensor* context, const sdh_tensor* y, const char* path);

/* Export the graph the reference builds for these inputs (no compute) for a host-side interpreter: <prefix>.json holds one record per
 * tensor (op, type, ne, nb, op_params, source ids, view root + byte offset) plus the node order; <prefix>.bin holds the data of every
 * leaf as float32 (weights read back from the backend buffer: F16 / BF16 / Q8_0 values are exact in f32; inputs from the host copies
 * the runner is about to upload).  oracle/graph_f64.py evaluates it in float64 -- the arbiter of the whole-model parity tests.
 * Returns the node count. */
int sdh_model_export_graph(sdh_model* m, const sdh_tensor* x, const sdh_tensor* timesteps,
                           const sdh_tensor* context, const sdh_tensor* y, const char* path_prefix);

/* w += value for the first parameter whose name contains `name_substr` and has `n_dims` dimensions, computed as a GRAPH on the model's
 * backend (ggml_add_inplace into the resident weight: the reference's LoRA apply, src/lora.hpp:934-937).  Returns the element count. */
int sdh_model_add_to_weight(sdh_model* m, const char* name_substr, int n_dims, float value);

/* Algorithmic FLOPs of the last graph built (sum over MUL_MAT / FLASH_ATTN_EXT / CONV_2D nodes,
 * SURVEY.md 8d), and its node counts. */
double sdh_model_last_graph_flops(const sdh_model* m);
int    sdh_model_last_graph_nodes(const sdh_model* m);

/* Reference sampler: the reference's sample_k_diffusion (denoiser.hpp:2794) with the reference's
 * CompVis denoiser / discrete schedule, CFG combine as in stable-diffusion.cpp:2855-2876.
 *   method: "euler_a" | "euler"    steps: e.g. 20    cfg_scale: e.g. 7.0   eta: ancestral eta (1.0)
 *   noise [W,H,C,1]: initial x_T noise (unit variance); cond/uncond: contexts [C_ctx,77,1]
 *   out: denoised latent like noise.  sigmas_out (optional, steps+1 floats) and
 *   timesteps_out (optional, steps floats) receive the scheduler's values for bit-exactness checks.
 *   n_forwards (optional) receives the number of model forwards issued.                         */
int sdh_sample(sdh_model* m, const char* method, int steps, float cfg_scale, float eta,
               uint64_t sampler_seed, const sdh_tensor* noise, const sdh_tensor* cond,
               const sdh_tensor* uncond, const sdh_tensor* y_cond, const sdh_tensor* y_uncond,
               sdh_tensor* out, float* sigmas_out, float* timesteps_out, int* n_forwards,
               double* wall_ms);

/* The reference's VAE::decode entry (src/model/vae/vae.hpp:171-221) on a vae_decoder model: tile_size > 0 enables its host-side tiling
 * (latent tiles of tile_size x tile_size with `overlap` in [0, 0.5], one graph_compute per tile, feathered merge on the host;
 * SURVEY.md 8a row a16), tile_size <= 0 decodes in one piece.  out: [8W, 8H, 3, N] scaled to [0, 1] like the reference. */
int sdh_vae_decode(sdh_model* m, const sdh_tensor* z, int tile_size, float overlap, sdh_tensor* out, double* wall_ms);

/* Builds the model's graph for these inputs (nothing is computed) and asks `fn` -- e.g. ggml_backend_b200_op_supported of the plugin -- about
 * every node: returns how many nodes `fn` rejects (0 = the whole graph runs on that backend, no ggml_backend_sched CPU fallback), or < 0
 * on error; the first rejected node is described in `first_unsupported`. */
typedef int (*sdh_op_supported_fn)(const struct ggml_tensor* op);
int sdh_model_check_ops(sdh_model* m, const sdh_tensor* x, const sdh_tensor* timesteps, const sdh_tensor* context, const sdh_tensor* y,
                        sdh_op_supported_fn fn, char* first_unsupported, size_t len);

/* CFG-batch split over a pair of GPUs: like sdh_sample, but this process evaluates only ONE branch per
 * step (role 0 = cond, 1 = uncond) and calls `exchange(mine, cond_out, uncond_out, n, user)` so the caller can
 * all-gather the eps prediction over NCCL; it must fill both outputs (n floats each) and return 0.  role < 0 or
 * exchange == NULL degenerates to sdh_sample.
 * role == 2: batched CFG on ONE device -- cond and uncond evaluated as a single N = 2 forward (x [W,H,C,2], context
 * [C,L,2], timesteps [2]); no exchange callback is used; n_forwards counts one forward per step. */
typedef int (*sdh_exchange_fn)(const float* mine, float* cond_out, float* uncond_out, size_t n, void* user);
int sdh_sample_split(sdh_model* m, const char* method, int steps, float cfg_scale, float eta,
                     uint64_t sampler_seed, const sdh_tensor* noise, const sdh_tensor* cond,
                     const sdh_tensor* uncond, const sdh_tensor* y_cond, const sdh_tensor* y_uncond,
                     sdh_tensor* out, float* sigmas_out, float* timesteps_out, int* n_forwards,
                     double* wall_ms, int role, sdh_exchange_fn exchange, void* user);

/* CFG split with the DEVICE-SIDE exchange (include/ggml-b200.h ggml_backend_b200_peer_mailbox_*): create returns this rank's 64-byte IPC
 * handle, connect takes the other rank's (NULL = loopback self test).  While connected, sdh_sample_split with role 0 / 1 and
 * exchange == NULL evaluates one branch per step and takes the other branch's eps prediction from the mailbox: no callback, no NCCL,
 * no host staging. */
int  sdh_model_mailbox_create(sdh_model* m, size_t bytes, void* handle_out64);
int  sdh_model_mailbox_connect(sdh_model* m, const void* peer_handle64);
void sdh_model_mailbox_close(sdh_model* m);

/* Counters of the model's backend instance when it is a B200 backend (include/ggml-b200.h ggml_b200_stats), as doubles:
 * [0] graphs [1] kernel_launches [2] nodes_executed [3] fused_nodes [4] last_graph_ms [5] total_graph_ms
 * [6] tc_gemm_launches [7..14] reserved[0..7], [15] unused, [16..31] ext[0..15] (see ggml-b200.h).  Returns <0 for other backends. */
int sdh_model_backend_stats(sdh_model* m, double* out, int n);
/* ggml_backend_b200_set_option on the model's backend (e.g. "fusion", "tc_gemm", "kernel_timing"). */
int sdh_model_set_backend_option(sdh_model* m, const char* key, int value);

/* Scheduler-only (no model): fills sigmas[steps+1] and t[steps] = sigma_to_t(sigmas[i]). */
int sdh_schedule(int steps, float* sigmas, float* timesteps);

/* Fill `n` floats with the reference's Philox N(0,1) stream for `seed` (core/rng_philox.hpp:100). */
int sdh_randn(uint64_t seed, float* dst, size_t n);

/* Run ONE ggml op (or the small op group the reference's wrapper emits) on `device` -- the same graph the
 * reference builds, so "CPU" gives the oracle's answer and "B200_0" ours.  Inputs are given as float32 and
 * stored on the device in the ggml type itypes[i] (0 = F32, 1 = F16, 30 = BF16: ggml_type values).
 *   op            inputs                       ip[]                               fp[]
 *   mul_mat       w [K,M,..], x [K,N,..]       -                                  -
 *   conv_2d       w [KW,KH,IC,OC], x, (bias)   s0,s1,p0,p1,d0,d1                  -      (ggml_conv_2d + bias add, ggml_extend.hpp:1131)
 *   im2col        w, x                         s0,s1,p0,p1,d0,d1,dst_type         -
 *   group_norm    x, (w [1,1,C,1], b), -       n_groups, silu(0/1)                eps    (ggml_extend.hpp:1502 [+ SiLU])
 *   norm|rms_norm x                            -                                  eps
 *   soft_max      x, (mask)                    -                                  scale, max_bias
 *   flash_attn    q, k, v, (mask)              -                                  scale  (k, v, mask stored F16)
 *   attention     q [C,Lq,N], k, v [C,Lk,N]    n_head, flash(0/1)                 -      (ggml_ext_attention_ext, :1349)
 *   upscale       x                            factor, mode                       -
 *   timestep_embedding  t [N]                  dim, max_period                    -
 *   unary         x                            ggml_unary_op                      -
 *   add|mul       a, b                         -                                  -
 *   scale         x                            -                                  scale, bias
 *   concat        a, b                         dim                                -
 *   cont_permute  x                            ax0,ax1,ax2,ax3                    -
 *   cpy           x                            dst_type                           -      (result read back as F32)
 * out->ne is filled; out->data (if non-NULL) must have room for the result (call once with data=NULL to size). */
int sdh_run_op(const char* device, const char* op, int n_in, const sdh_tensor* in, const int32_t* itypes,
               const int32_t* ip, const float* fp, sdh_tensor* out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* SD_B200_HARNESS_H */
