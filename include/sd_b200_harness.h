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

/* arch:    "sd15_unet" | "sdxl_unet" | "vae_decoder" | "flux_schnell" | "flux_tiny" | "unet_tiny"
 * wtype:   "f32" | "f16" | "bf16" | "q8_0"  (dtype of Linear weights; conv weights are always F16 as
 *           in the reference, ggml_extend.hpp:3600-3603; norm/bias are F32)
 * flags:   bit0 = flash attention graph variant (--diffusion-fa), bit1 = conv2d direct
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

/* Scheduler-only (no model): fills sigmas[steps+1] and t[steps] = sigma_to_t(sigmas[i]). */
int sdh_schedule(int steps, float* sigmas, float* timesteps);

/* Fill `n` floats with the reference's Philox N(0,1) stream for `seed` (core/rng_philox.hpp:100). */
int sdh_randn(uint64_t seed, float* dst, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* SD_B200_HARNESS_H */
