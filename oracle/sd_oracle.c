/* sd_oracle.c -- TEST INFRASTRUCTURE.  Plain-C restatement of the reference's host-side scheduler /
 * sampler index math (the part of the hot path that must be BIT-EXACT), written against libm exactly as
 * the reference's C++ is (std::log/exp/sqrt on float == logf/expf/sqrtf).  Only tests/, smoke() and
 * bench.py's CPU-baseline legs may load this; the product never does.
 *
 * Pinned by tests/test_oracle.py against (a) the reference's own code compiled into host/_ref/libsd_harness.so
 * (sdh_schedule) and (b) the committed golden fixture tests/golden/schedule_sd15_20.json generated from it.
 *
 *   sd_alphas_cumprod      src/stable-diffusion.cpp:173-186   (calculate_alphas_cumprod)
 *   sd_compvis_tables      src/stable-diffusion.cpp:666-681   (refresh_compvis_denoiser_sigmas)
 *   sd_t_to_sigma          src/runtime/denoiser.hpp:1166-1172 (CompVisDenoiser::t_to_sigma)
 *   sd_sigma_to_t          src/runtime/denoiser.hpp:1140-1164 (CompVisDenoiser::sigma_to_t)
 *   sd_discrete_sigmas     src/runtime/denoiser.hpp:32-54     (DiscreteScheduler::get_sigmas)
 *   sd_scalings            src/runtime/denoiser.hpp:1174-1179 (CompVisDenoiser::get_scalings)
 *   sd_ancestral_step      src/runtime/denoiser.hpp:1447-1467 (get_ancestral_step)
 */
#include <math.h>
#include <stdint.h>

#define SD_TIMESTEPS 1000

void sd_alphas_cumprod(float* alphas_cumprod) {
    const float linear_start = 0.00085f, linear_end = 0.0120f;
    float ls_sqrt = sqrtf(linear_start);
    float le_sqrt = sqrtf(linear_end);
    float amount  = le_sqrt - ls_sqrt;
    float product = 1.0f;
    for (int i = 0; i < SD_TIMESTEPS; i++) {
        float beta = ls_sqrt + amount * ((float)i / (SD_TIMESTEPS - 1));
        product *= 1.0f - powf(beta, 2.0f);
        alphas_cumprod[i] = product;
    }
}

void sd_compvis_tables(float* sigmas, float* log_sigmas) {
    float ac[SD_TIMESTEPS];
    sd_alphas_cumprod(ac);
    for (int i = 0; i < SD_TIMESTEPS; i++) {
        sigmas[i]     = sqrtf((1 - ac[i]) / ac[i]);
        log_sigmas[i] = logf(sigmas[i]);
    }
}

float sd_t_to_sigma(const float* log_sigmas, float t) {
    int low_idx     = (int)floorf(t);
    int high_idx    = (int)ceilf(t);
    float w         = t - (float)low_idx;
    float log_sigma = (1.0f - w) * log_sigmas[low_idx] + w * log_sigmas[high_idx];
    return expf(log_sigma);
}

float sd_sigma_to_t(const float* log_sigmas, float sigma) {
    float log_sigma = logf(sigma);
    int low_idx = 0;
    for (int i = 0; i < SD_TIMESTEPS; i++) {
        float dist = log_sigma - log_sigmas[i];
        if (dist >= 0) low_idx++;
    }
    low_idx = low_idx - 1;
    if (low_idx < 0) low_idx = 0;
    if (low_idx > SD_TIMESTEPS - 2) low_idx = SD_TIMESTEPS - 2;
    int high_idx = low_idx + 1;
    float low  = log_sigmas[low_idx];
    float high = log_sigmas[high_idx];
    float w    = (low - log_sigma) / (low - high);
    w          = fmaxf(0.f, fminf(1.f, w));
    return (1.0f - w) * low_idx + w * high_idx;
}

/* fills n+1 sigmas; returns count written */
int sd_discrete_sigmas(const float* log_sigmas, uint32_t n, float* out) {
    int t_max = SD_TIMESTEPS - 1;
    if (n == 0) return 0;
    if (n == 1) {
        out[0] = sd_t_to_sigma(log_sigmas, (float)t_max);
        out[1] = 0;
        return 2;
    }
    float step = (float)t_max / (float)(n - 1);
    for (uint32_t i = 0; i < n; ++i) {
        float t = t_max - step * i;
        out[i]  = sd_t_to_sigma(log_sigmas, t);
    }
    out[n] = 0;
    return (int)n + 1;
}

void sd_scalings(float sigma, float* c_skip, float* c_out, float* c_in) {
    const float sigma_data = 1.0f;
    *c_skip = 1.0f;
    *c_out  = -sigma;
    *c_in   = 1.0f / sqrtf(sigma * sigma + sigma_data * sigma_data);
}

void sd_ancestral_step(float sigma_from, float sigma_to, float eta, float* sigma_down_out, float* sigma_up_out) {
    float sigma_up   = 0.0f;
    float sigma_down = sigma_to;
    if (eta > 0.0f) {
        float sigma_from_sq = sigma_from * sigma_from;
        float sigma_to_sq   = sigma_to * sigma_to;
        if (sigma_from_sq > 0.0f) {
            float term = sigma_to_sq * (sigma_from_sq - sigma_to_sq) / sigma_from_sq;
            sigma_up   = fminf(sigma_to, eta * sqrtf(fmaxf(term, 0.0f)));
        }
        float sigma_down_sq = sigma_to_sq - sigma_up * sigma_up;
        sigma_down          = sigma_down_sq > 0.0f ? sqrtf(sigma_down_sq) : 0.0f;
    }
    *sigma_down_out = sigma_down;
    *sigma_up_out   = sigma_up;
}

/* convenience: the whole 20-step style schedule in one call: sigmas[n+1], timesteps[n] */
void sd_schedule(uint32_t n, float* sigmas, float* timesteps) {
    float sg[SD_TIMESTEPS], ls[SD_TIMESTEPS];
    sd_compvis_tables(sg, ls);
    sd_discrete_sigmas(ls, n, sigmas);
    for (uint32_t i = 0; i < n; ++i) timesteps[i] = sd_sigma_to_t(ls, sigmas[i]);
}
