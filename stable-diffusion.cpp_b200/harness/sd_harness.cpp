// sd_harness.cpp -- synthetic-weight whole-model harness (C-ABI in include/sd_b200_harness.h).
//
// HOST code only.  It compiles the reference's own header-only graph builders and sampler
// (included from /root/reference at build time, nothing is copied) and runs them on whichever
// ggml backend the caller names: "CPU" (the oracle) or "B200_<i>" (this repo's plugin).
//   UNetModelRunner            src/model/diffusion/unet.hpp:747-858
//   AutoEncoderKL              src/model/vae/auto_encoder_kl.hpp:662-760
//   Flux::FluxRunner           src/model/diffusion/flux.hpp
//   sample_k_diffusion         src/runtime/denoiser.hpp:2794
//   ClassifierFreeGuidance     src/runtime/guidance.cpp:149-178
// Our own code here is: the synthetic RunnerWeightManager (weight_manager.h:10-17 interface),
// seeded weight generation, the denoise callback (a minimal restatement of the lambda at
// src/stable-diffusion.cpp:2625-2895 for the plain cond/uncond CFG case), and the C-ABI glue.

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <sched.h>
#include <vector>

#include "ggml-backend.h"
#include "ggml.h"

#include "model/diffusion/flux.hpp"
#include "model/diffusion/mmdit.hpp"
#include "model/diffusion/wan.hpp"
#include "model/te/clip.hpp"
#include "model/te/t5.hpp"
#include "model/diffusion/unet.hpp"
#include "model/vae/auto_encoder_kl.hpp"
#include "model/vae/wan_vae.hpp"
#include "runtime/denoiser.hpp"
#include "runtime/guidance.h"
#include "core/rng_philox.hpp"

#include "../../include/sd_b200_harness.h"

namespace {

thread_local std::string g_last_error;

int fail(const std::string& msg) {
    g_last_error = msg;
    fprintf(stderr, "[sd_harness] error: %s\n", msg.c_str());
    return -1;
}

// ---------------------------------------------------------------- seeded weight generation
inline uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t fnv1a(const std::string& s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ull;
    }
    return h;
}
// uniform in [-1, 1)
inline float u11(uint64_t& s) {
    return (float)((double)(splitmix64(s) >> 11) * (2.0 / 9007199254740992.0) - 1.0);
}

// Fill `n` floats: value = center + amp * U(-1,1); chunked so the stream does not depend on the
// number of threads used.
void fill_uniform(float* dst, int64_t n, uint64_t seed, float center, float amp) {
    const int64_t chunk = 1 << 16;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            int64_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            uint64_t s = seed ^ (0xD6E8FEB86659FD93ull * (uint64_t)(c + 1));
            int64_t b = c * chunk, e = std::min(n, b + chunk);
            for (int64_t i = b; i < e; ++i) dst[i] = center + amp * u11(s);
        }
    };
    if (nchunks < 4) { work(); return; }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nt; ++i) th.emplace_back(work);
    for (auto& t : th) t.join();
}

// rows [0, n) in blocks over up to 16 host threads (weight type conversion of multi-GB synthetic checkpoints)
template <typename F>
void parallel_rows(int64_t n, int64_t block, F fn) {
    const int64_t nblocks = (n + block - 1) / block;
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            fn(b * block, std::min(n, (b + 1) * block));
        }
    };
    if (nblocks < 4) { work(); return; }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nt; ++i) th.emplace_back(work);
    for (auto& t : th) t.join();
}

// ---------------------------------------------------------------- synthetic weight manager
// All parameters live in ONE buffer on the compute backend for the whole life of the model
// (the reference keeps weights resident too: free_compute_params=false, unet.hpp:838).
struct SyntheticWeights : public RunnerWeightManager {
    ggml_backend_buffer_t buffer = nullptr;
    size_t bytes                 = 0;
    int count                    = 0;

    ~SyntheticWeights() override {
        if (buffer) ggml_backend_buffer_free(buffer);
    }
    bool assign_compute_backend(const std::vector<ggml_tensor*>&, ggml_backend_t) override { return true; }
    bool prepare_params(const std::vector<ggml_tensor*>&) override { return true; }
    void release_compute_backend_params(const std::vector<ggml_tensor*>&) override {}
    void release_params_backend_params(const std::vector<ggml_tensor*>&) override {}

    // fill == false: tensors are placed in the buffer but keep whatever the allocation holds (graph / claim walks that never run the model)
    bool materialise(ggml_backend_t backend, const std::map<std::string, ggml_tensor*>& tensors, uint64_t seed, bool fill = true) {
        ggml_backend_buffer_type_t buft = ggml_backend_get_default_buffer_type(backend);
        size_t align = ggml_backend_buft_get_alignment(buft);
        size_t total = 0;
        for (auto& kv : tensors) {
            size_t sz = ggml_backend_buft_get_alloc_size(buft, kv.second);
            total += (sz + align - 1) / align * align;
        }
        buffer = ggml_backend_buft_alloc_buffer(buft, total + align);
        if (!buffer) return false;
        ggml_backend_buffer_set_usage(buffer, GGML_BACKEND_BUFFER_USAGE_WEIGHTS);
        char* base = (char*)ggml_backend_buffer_get_base(buffer);
        size_t off = 0;
        std::vector<float> tmp;
        std::vector<uint8_t> conv;
        for (auto& kv : tensors) {
            ggml_tensor* t = kv.second;
            size_t sz = ggml_backend_buft_get_alloc_size(buft, t);
            if (ggml_backend_tensor_alloc(buffer, t, base + off) != GGML_STATUS_SUCCESS) return false;
            off += (sz + align - 1) / align * align;
            if (!fill) { bytes += ggml_nbytes(t); count++; continue; }

            const std::string& name = kv.first;
            int64_t n = ggml_nelements(t);
            tmp.resize(n);
            uint64_t s = seed ^ fnv1a(name);
            bool is_bias = name.size() >= 4 && name.compare(name.size() - 4, 4, "bias") == 0;
            int nd = ggml_n_dims(t);
            if (nd <= 1) {
                // norm scales ~ 1 +- 0.1 ; biases (and other vectors) ~ +-0.05
                bool is_scale = !is_bias && (name.find("norm") != std::string::npos || name.find("scale") != std::string::npos ||
                                             name.find("ln_") != std::string::npos || name.find(".0.weight") != std::string::npos ||
                                             name.find("weight") != std::string::npos || name.find("gamma") != std::string::npos);
                if (is_scale) fill_uniform(tmp.data(), n, s, 1.0f, 0.1f);
                else fill_uniform(tmp.data(), n, s, 0.0f, 0.05f);
            } else {
                // Linear [in,out] / Conv [KW,KH,IC,OC]: fan_in = prod(ne[0..nd-2]); U(+-sqrt(3/fan_in)) has variance 1/fan_in
                int64_t fan_in = 1;
                for (int d = 0; d < nd - 1; ++d) fan_in *= t->ne[d];
                fill_uniform(tmp.data(), n, s, 0.0f, std::sqrt(3.0f / (float)fan_in));
            }
            if (t->type == GGML_TYPE_F32) {
                ggml_backend_tensor_set(t, tmp.data(), 0, ggml_nbytes(t));
            } else {
                conv.resize(ggml_nbytes(t));
                const ggml_type_traits* tt = ggml_get_type_traits(t->type);
                if (!tt->from_float_ref) return false;
                // quantize row by row (rows are ne[0] long)
                int64_t nrows = n / t->ne[0];
                size_t row_bytes = ggml_row_size(t->type, t->ne[0]);
                const int64_t ne0 = t->ne[0];
                const float* srcp = tmp.data();
                uint8_t* dstp = conv.data();
                parallel_rows(nrows, std::max<int64_t>(1, (1 << 18) / std::max<int64_t>(ne0, 1)), [&](int64_t r0, int64_t r1) {
                    for (int64_t r = r0; r < r1; ++r) tt->from_float_ref(srcp + r * ne0, dstp + r * row_bytes, ne0);
                });
                ggml_backend_tensor_set(t, conv.data(), 0, ggml_nbytes(t));
            }
            bytes += ggml_nbytes(t);
            count++;
        }
        return true;
    }
};

enum Arch { ARCH_UNET, ARCH_VAE, ARCH_FLUX, ARCH_MMDIT, ARCH_WAN, ARCH_CLIP, ARCH_WANVAE, ARCH_T5 };

ggml_type parse_wtype(const char* w) {
    std::string s = w ? w : "f32";
    if (s == "f16") return GGML_TYPE_F16;
    if (s == "bf16") return GGML_TYPE_BF16;
    if (s == "q8_0") return GGML_TYPE_Q8_0;
    return GGML_TYPE_F32;
}

// graph statistics: algorithmic FLOPs per SURVEY.md 8(d)
double graph_flops(ggml_cgraph* gf) {
    double fl = 0;
    int n = ggml_graph_n_nodes(gf);
    for (int i = 0; i < n; ++i) {
        ggml_tensor* t = ggml_graph_node(gf, i);
        if (t->op == GGML_OP_MUL_MAT) {
            fl += 2.0 * (double)t->src[0]->ne[0] * (double)t->ne[0] * (double)t->ne[1] * (double)t->ne[2] * (double)t->ne[3];
        } else if (t->op == GGML_OP_FLASH_ATTN_EXT) {
            const ggml_tensor* q = t->src[0]; const ggml_tensor* k = t->src[1]; const ggml_tensor* v = t->src[2];
            // q [d, Lq, h, N], k [d, Lk, h_kv, N], v [dv, Lk, h_kv, N]
            fl += 2.0 * (double)q->ne[1] * (double)k->ne[1] * (double)q->ne[2] * (double)q->ne[3] * ((double)q->ne[0] + (double)v->ne[0]);
        } else if (t->op == GGML_OP_CONV_2D) {
            const ggml_tensor* w = t->src[0];
            fl += 2.0 * (double)w->ne[0] * w->ne[1] * w->ne[2] * (double)t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3];
        }
    }
    return fl;
}

}  // namespace

struct sdh_model {
    Arch arch;
    std::string device;
    ggml_backend_t backend = nullptr;
    std::shared_ptr<SyntheticWeights> weights;
    std::unique_ptr<UNetModelRunner> unet;
    std::unique_ptr<AutoEncoderKL> vae;
    std::unique_ptr<Flux::FluxRunner> flux;
    std::unique_ptr<MMDiTRunner> mmdit;
    std::unique_ptr<WAN::WanRunner> wan;
    std::unique_ptr<CLIPTextModelRunner> clip;
    std::unique_ptr<WAN::WanVAERunner> wan_vae;
    std::unique_ptr<T5Runner> t5;
    SDVersion version = VERSION_SD1;
    int n_threads     = 1;
    double last_flops = 0;
    int last_nodes    = 0;
    std::map<std::string, ggml_tensor*> params;   // parameter tensors by checkpoint name (resident on the backend)
    bool mailbox = false;                          // the backend's peer mailbox is connected (CFG split over NVLink)
};

namespace {

sd::Tensor<float> to_sd(const sdh_tensor* t) {
    if (!t || !t->data) return {};
    std::vector<int64_t> shape;
    int nd = 4;
    while (nd > 1 && t->ne[nd - 1] == 1) nd--;
    for (int i = 0; i < nd; ++i) shape.push_back(t->ne[i]);
    int64_t n = 1;
    for (auto v : shape) n *= v;
    std::vector<float> data(t->data, t->data + n);
    return sd::Tensor<float>(shape, std::move(data));
}

// keeps all 4 dims (x must stay 4-D even when N == 1)
sd::Tensor<float> to_sd_nd(const sdh_tensor* t, int nd) {
    if (!t || !t->data) return {};
    std::vector<int64_t> shape(t->ne, t->ne + nd);
    int64_t n = 1;
    for (auto v : shape) n *= v;
    std::vector<float> data(t->data, t->data + n);
    return sd::Tensor<float>(shape, std::move(data));
}

int from_sd(const sd::Tensor<float>& s, sdh_tensor* out) {
    if (s.empty()) return fail("model returned an empty tensor (graph_compute failed)");
    const auto& sh = s.shape();
    for (int i = 0; i < 4; ++i) out->ne[i] = i < (int)sh.size() ? sh[i] : 1;
    if (out->data) memcpy(out->data, s.data(), sizeof(float) * s.numel());
    return 0;
}

template <class BuildFn>
std::map<std::string, ggml_type> pass1_types(BuildFn build, ggml_type wtype) {
    // first pass with an empty storage map: discover parameter names / shapes, then ask for
    // `wtype` on every >=2-D F32 weight (Linear weights; conv weights are already F16)
    return {};
}

String2TensorStorage make_storage_map(const std::map<std::string, ggml_tensor*>& tensors, ggml_type wtype) {
    String2TensorStorage m;
    if (wtype == GGML_TYPE_F32) return m;
    for (auto& kv : tensors) {
        ggml_tensor* t = kv.second;
        if (ggml_n_dims(t) != 2 || t->type != GGML_TYPE_F32) continue;
        if (t->ne[0] % ggml_blck_size(wtype) != 0) continue;
        TensorStorage ts(kv.first, wtype, t->ne, 2, 0);
        m[kv.first] = ts;
    }
    return m;
}

// default CPU thread count: affinity mask and cgroup quota aware, capped at 16 (a ggml thread pool larger than the
// cores the container may really use spins itself ~30x slower; callers that want the maximum pass n_threads explicitly)
int default_cpu_threads() {
    if (const char* e = getenv("SDH_CPU_THREADS")) {
        int v = atoi(e);
        if (v > 0) return v;
    }
    int n = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min(n, CPU_COUNT(&set));
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char q[64];
        long period = 0;
        if (fscanf(f, "%63s %ld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) n = std::min(n, std::max(1, (int)(atol(q) / period)));
        fclose(f);
    }
    return std::max(1, std::min(n, 16));
}

ggml_backend_t init_device(const std::string& device, int n_threads) {
    if (n_threads <= 0) n_threads = default_cpu_threads();
    ggml_backend_dev_t dev = ggml_backend_dev_by_name(device.c_str());
    if (!dev) return nullptr;
    ggml_backend_t be = ggml_backend_dev_init(dev, nullptr);
    if (!be) return nullptr;
    if (ggml_backend_dev_type(dev) == GGML_BACKEND_DEVICE_TYPE_CPU) {
        auto reg = ggml_backend_dev_backend_reg(dev);
        auto fn  = (ggml_backend_set_n_threads_t)ggml_backend_reg_get_proc_address(reg, "ggml_backend_set_n_threads");
        if (fn) fn(be, n_threads);
    }
    return be;
}

}  // namespace

extern "C" {

const char* sdh_last_error(void) { return g_last_error.c_str(); }

int sdh_load_backend(const char* so_path) {
    ggml_backend_reg_t reg = ggml_backend_load(so_path);
    if (!reg) return fail(std::string("ggml_backend_load failed for ") + so_path);
    return (int)ggml_backend_dev_count();
}

int sdh_device_count(void) { return (int)ggml_backend_dev_count(); }

const char* sdh_device_name(int index) {
    if (index < 0 || index >= (int)ggml_backend_dev_count()) return nullptr;
    return ggml_backend_dev_name(ggml_backend_dev_get(index));
}

sdh_model* sdh_model_create(const char* device, const char* arch, const char* wtype_s, int flags, uint64_t seed, int n_threads) {
    std::string a = arch ? arch : "";
    ggml_type wtype = parse_wtype(wtype_s);
    auto m = std::make_unique<sdh_model>();
    m->device    = device ? device : "CPU";
    m->n_threads = n_threads > 0 ? n_threads : default_cpu_threads();
    m->backend   = init_device(m->device, m->n_threads);
    if (!m->backend) { fail("no such device: " + m->device); return nullptr; }
    m->weights = std::make_shared<SyntheticWeights>();
    bool fa = flags & 1, direct = flags & 2;

    std::map<std::string, ggml_tensor*> tensors;
    if (a == "sd15_unet" || a == "sdxl_unet" || a == "unet_tiny") {
        m->arch    = ARCH_UNET;
        m->version = a == "sdxl_unet" ? VERSION_SDXL : (a == "unet_tiny" ? VERSION_SD1_TINY_UNET : VERSION_SD1);
        const std::string prefix = "model.diffusion_model";
        String2TensorStorage smap;
        {
            UNetModelRunner probe(m->backend, {}, prefix, m->version, m->weights);
            std::map<std::string, ggml_tensor*> pt;
            probe.get_param_tensors(pt, prefix);
            smap = make_storage_map(pt, wtype);
        }
        m->unet = std::make_unique<UNetModelRunner>(m->backend, smap, prefix, m->version, m->weights);
        m->unet->get_param_tensors(tensors, prefix);
        m->unet->set_flash_attention_enabled(fa);
        m->unet->set_conv2d_direct_enabled(direct);
    } else if (a == "vae_decoder" || a == "vae_decoder_sdxl") {
        m->arch    = ARCH_VAE;
        m->version = a == "vae_decoder_sdxl" ? VERSION_SDXL : VERSION_SD1;
        const std::string prefix = "first_stage_model";
        String2TensorStorage smap;
        {
            AutoEncoderKL probe(m->backend, {}, prefix, true, false, m->version, m->weights);
            std::map<std::string, ggml_tensor*> pt;
            probe.get_param_tensors(pt);
            smap = make_storage_map(pt, wtype);
        }
        m->vae = std::make_unique<AutoEncoderKL>(m->backend, smap, prefix, true, false, m->version, m->weights);
        m->vae->get_param_tensors(tensors);
        m->vae->set_flash_attention_enabled(fa);
        m->vae->set_conv2d_direct_enabled(direct);
    } else if (a == "flux_schnell" || a == "flux_tiny" || a == "flux_1x1") {
        m->arch    = ARCH_FLUX;
        m->version = VERSION_FLUX;
        const std::string prefix = "model.diffusion_model";
        String2TensorStorage smap;
        {
            // depth is detected from weight names (flux.hpp detect_from_weights): FLUX.1 has 19 double + 38 single blocks, the tiny
            // variant declares 2 + 2
            // flux_1x1: ONE double-stream + ONE single-stream block at the full FLUX.1 width (hidden 3072, 24 heads): full-size block parity
            const int n_double = a == "flux_tiny" ? 2 : (a == "flux_1x1" ? 1 : 19), n_single = a == "flux_tiny" ? 2 : (a == "flux_1x1" ? 1 : 38);
            int64_t ne2[2] = {3072, 3072};
            for (int i = 0; i < n_double; ++i) {
                std::string n1 = prefix + ".double_blocks." + std::to_string(i) + ".img_attn.proj.weight";
                smap[n1] = TensorStorage(n1, wtype, ne2, 2, 0);
            }
            for (int i = 0; i < n_single; ++i) {
                std::string n2 = prefix + ".single_blocks." + std::to_string(i) + ".modulation.lin.weight";
                smap[n2] = TensorStorage(n2, wtype, ne2, 2, 0);
            }
        }
        {
            Flux::FluxRunner probe(m->backend, smap, prefix, m->version, m->weights);
            std::map<std::string, ggml_tensor*> pt;
            probe.get_param_tensors(pt, prefix);
            String2TensorStorage typed = make_storage_map(pt, wtype);
            for (auto& kv : typed) smap[kv.first] = kv.second;
        }
        m->flux = std::make_unique<Flux::FluxRunner>(m->backend, smap, prefix, m->version, m->weights);
        m->flux->get_param_tensors(tensors, prefix);
        m->flux->set_flash_attention_enabled(fa);
    } else if (a == "mmdit_sd3") {
        // SD3-medium: the MMDiTConfig defaults (depth 24, hidden 1536, 2 B parameters; mmdit.hpp:16-31)
        m->arch    = ARCH_MMDIT;
        m->version = VERSION_SD3;
        const std::string prefix = "model.diffusion_model";
        String2TensorStorage smap;
        {
            MMDiTRunner probe(m->backend, {}, prefix, m->weights);
            std::map<std::string, ggml_tensor*> pt;
            probe.get_param_tensors(pt, prefix);
            smap = make_storage_map(pt, wtype);
        }
        m->mmdit = std::make_unique<MMDiTRunner>(m->backend, smap, prefix, m->weights);
        m->mmdit->get_param_tensors(tensors, prefix);
        m->mmdit->set_flash_attention_enabled(fa);
    } else if (a == "wan_1_3b") {
        // Wan2.1-T2V-1.3B: 30 blocks of dim 1536 / 12 heads (wan.hpp:808-837); depth is detected from weight names
        m->arch    = ARCH_WAN;
        m->version = VERSION_WAN2;
        const std::string prefix = "model.diffusion_model";
        String2TensorStorage smap;
        int64_t ne2[2] = {1536, 1536};
        for (int i = 0; i < 30; ++i) {
            std::string n1 = prefix + ".blocks." + std::to_string(i) + ".self_attn.q.weight";
            smap[n1] = TensorStorage(n1, wtype, ne2, 2, 0);
        }
        {
            WAN::WanRunner probe(m->backend, smap, prefix, m->version, m->weights);
            std::map<std::string, ggml_tensor*> pt;
            probe.get_param_tensors(pt, prefix);
            String2TensorStorage typed = make_storage_map(pt, wtype);
            for (auto& kv : typed) smap[kv.first] = kv.second;
        }
        m->wan = std::make_unique<WAN::WanRunner>(m->backend, smap, prefix, m->version, m->weights);
        m->wan->get_param_tensors(tensors, prefix);
        m->wan->set_flash_attention_enabled(fa);
    } else if (a == "wan_vae_decoder") {
        // Wan2.1 VAE decoder (causal 3-D convolutions, src/model/vae/wan_vae.hpp): latent [W, H, T, 16] -> video [8W, 8H, 4(T-1)+1, 3]
        m->arch    = ARCH_WANVAE;
        m->version = VERSION_WAN2;
        const std::string prefix = "first_stage_model";
        String2TensorStorage smap;
        {
            WAN::WanVAERunner probe(m->backend, {}, prefix, true, m->version, m->weights);
            std::map<std::string, ggml_tensor*> pt;
            probe.get_param_tensors(pt);
            smap = make_storage_map(pt, wtype);
        }
        m->wan_vae = std::make_unique<WAN::WanVAERunner>(m->backend, smap, prefix, true, m->version, m->weights);
        m->wan_vae->get_param_tensors(tensors);
    } else if (a == "clip_l") {
        // the CLIP ViT-L/14 text encoder of SD1.x / SDXL (src/model/te/clip.hpp), the stage right before the hot path
        m->arch    = ARCH_CLIP;
        m->version = VERSION_SD1;
        const std::string prefix = "cond_stage_model.transformer.text_model";
        String2TensorStorage smap;
        {
            CLIPTextModelRunner probe(m->backend, {}, prefix, OPENAI_CLIP_VIT_L_14, true, false, m->weights);
            std::map<std::string, ggml_tensor*> pt;
            probe.get_param_tensors(pt, prefix);
            smap = make_storage_map(pt, wtype);
        }
        m->clip = std::make_unique<CLIPTextModelRunner>(m->backend, smap, prefix, OPENAI_CLIP_VIT_L_14, true, false, m->weights);
        m->clip->get_param_tensors(tensors, prefix);
        m->clip->set_flash_attention_enabled(fa);
    } else if (a == "t5_xxl_4l" || a == "t5_xxl") {
        // the T5-XXL encoder of Flux / SD3 (src/model/te/t5.hpp): RMS-style T5LayerNorm, relative-position-bias attention
        // (GET_ROWS of the bucket table, bias added to the scores), gated-GELU feed forward.  t5_xxl_4l: 4 of the 24 layers at full width.
        m->arch    = ARCH_T5;
        m->version = VERSION_FLUX;
        const std::string prefix = "text_encoders.t5xxl.transformer";
        const int n_layers = a == "t5_xxl_4l" ? 4 : 24;
        String2TensorStorage smap;
        int64_t ne2[2] = {4096, 4096};
        for (int i = 0; i < n_layers; ++i) {
            std::string n1 = prefix + ".encoder.block." + std::to_string(i) + ".layer.0.SelfAttention.q.weight";
            smap[n1] = TensorStorage(n1, wtype, ne2, 2, 0);
        }
        {
            T5Runner probe(m->backend, smap, prefix, false, m->weights);
            std::map<std::string, ggml_tensor*> pt;
            probe.get_param_tensors(pt, prefix);
            String2TensorStorage typed = make_storage_map(pt, wtype);
            for (auto& kv : typed) smap[kv.first] = kv.second;
        }
        m->t5 = std::make_unique<T5Runner>(m->backend, smap, prefix, false, m->weights);
        m->t5->get_param_tensors(tensors, prefix);
        m->t5->set_flash_attention_enabled(fa);
    } else {
        fail("unknown arch: " + a);
        return nullptr;
    }
    if (!m->weights->materialise(m->backend, tensors, seed, (flags & 4) == 0)) {
        fail("weight allocation failed");
        return nullptr;
    }
    m->params = tensors;
    return m.release();
}

void sdh_model_free(sdh_model* m) {
    if (!m) return;
    m->unet.reset();
    m->vae.reset();
    m->flux.reset();
    m->mmdit.reset();
    m->wan.reset();
    m->clip.reset();
    m->t5.reset();
    m->wan_vae.reset();
    m->weights.reset();
    if (m->backend) ggml_backend_free(m->backend);
    delete m;
}

size_t sdh_model_param_bytes(const sdh_model* m) { return m->weights ? m->weights->bytes : 0; }
int sdh_model_param_count(const sdh_model* m) { return m->weights ? m->weights->count : 0; }

int sdh_model_out_shape(sdh_model* m, const sdh_tensor* x, int64_t out_ne[4]) {
    if (!m || !x) return fail("null argument");
    for (int i = 0; i < 4; ++i) out_ne[i] = x->ne[i];
    if (m->arch == ARCH_T5) {        // ids [n_token, N] -> hidden states [4096, n_token, N]
        out_ne[0] = 4096; out_ne[1] = x->ne[0]; out_ne[2] = x->ne[1]; out_ne[3] = 1;
    }
    if (m->arch == ARCH_CLIP) {      // ids [n_token, N] -> hidden states [768, n_token, N]
        out_ne[0] = 768; out_ne[1] = x->ne[0]; out_ne[2] = x->ne[1]; out_ne[3] = 1;
    }
    if (m->arch == ARCH_WANVAE) {
        out_ne[0] = x->ne[0] * 8; out_ne[1] = x->ne[1] * 8; out_ne[2] = (x->ne[2] - 1) * 4 + 1; out_ne[3] = 3;
    }
    if (m->arch == ARCH_VAE) {
        out_ne[0] = x->ne[0] * 8;
        out_ne[1] = x->ne[1] * 8;
        out_ne[2] = 3;
    }
    return 0;
}

// text-encoder input: token ids arrive as floats in x ([n_token, N, 1, 1]); CLIP wants int32 [n_token, N]
static sd::Tensor<int32_t> token_ids(const sd::Tensor<float>& x) {
    std::vector<int32_t> ids(x.values().begin(), x.values().end());
    return sd::Tensor<int32_t>({x.shape()[0], x.shape()[1]}, std::move(ids));
}

static sd::Tensor<float> run_model(sdh_model* m, const sd::Tensor<float>& x, const sd::Tensor<float>& t,
                                   const sd::Tensor<float>& ctx, const sd::Tensor<float>& y) {
    switch (m->arch) {
        case ARCH_CLIP:
            return m->clip->compute(m->n_threads, token_ids(x), 0, nullptr, 0, false, -1, false, false, false);
        case ARCH_T5:
            return m->t5->compute(m->n_threads, token_ids(x), {}, false, false, false);
        case ARCH_WANVAE: {
            // a 4-D tensor would be taken for an image [W,H,C,N] (wan_vae.hpp:1384-1388): video latents go in as [W,H,T,C,1]
            std::vector<int64_t> shape5 = x.shape();
            shape5.push_back(1);
            sd::Tensor<float> x5(shape5, std::vector<float>(x.values()));
            sd::Tensor<float> r = m->wan_vae->_compute(m->n_threads, x5, true);
            if (r.empty()) return r;
            std::vector<int64_t> s4(r.shape().begin(), r.shape().begin() + 4);
            return sd::Tensor<float>(s4, std::vector<float>(r.values()));
        }
        case ARCH_UNET:
            return m->unet->compute(m->n_threads, x, t, ctx, {}, y);
        case ARCH_VAE:
            return m->vae->_compute(m->n_threads, x, true);
        case ARCH_FLUX: {
            sd::Tensor<float> guidance;  // schnell: no guidance embed
            return m->flux->compute(m->n_threads, x, t, ctx, {}, y, guidance);
        }
        case ARCH_MMDIT:
            return m->mmdit->compute(m->n_threads, x, t, ctx, y);
        case ARCH_WAN:
            return m->wan->compute(m->n_threads, x, t, ctx);
    }
    return {};
}

int sdh_model_forward(sdh_model* m, const sdh_tensor* x, const sdh_tensor* timesteps, const sdh_tensor* context,
                      const sdh_tensor* y, sdh_tensor* out, double* wall_ms) {
    if (!m || !x || !out) return fail("null argument");
    auto xs = to_sd_nd(x, 4);
    auto ts = to_sd_nd(timesteps, 1);
    auto cs = to_sd_nd(context, 3);
    auto ys = to_sd_nd(y, 2);
    auto t0 = std::chrono::steady_clock::now();
    sd::Tensor<float> r = run_model(m, xs, ts, cs, ys);
    auto t1 = std::chrono::steady_clock::now();
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    return from_sd(r, out);
}

static ggml_cgraph* build_only(sdh_model* m, const sd::Tensor<float>& x, const sd::Tensor<float>& t,
                               const sd::Tensor<float>& ctx, const sd::Tensor<float>& y);

// VAE::decode of the reference (src/model/vae/vae.hpp:171-221) as the txt2img path calls it: with tile_size > 0 the latent is split into
// overlapping tiles on the HOST (sd_tiling, ggml_extend.hpp:691-951), each tile is decoded by one graph_compute on the backend, and the
// tiles are feather-blended on the host; tile_size <= 0 decodes in one piece.  The output is the reference's [0,1]-scaled image.
int sdh_vae_decode(sdh_model* m, const sdh_tensor* z, int tile_size, float overlap, sdh_tensor* out, double* wall_ms) {
    if (!m || !z || !out || m->arch != ARCH_VAE) return fail("sdh_vae_decode needs a vae_decoder model");
    auto zs = to_sd_nd(z, 4);
    sd_tiling_params_t tp;
    memset(&tp, 0, sizeof(tp));
    tp.enabled        = tile_size > 0;
    tp.tile_size_x    = tile_size;
    tp.tile_size_y    = tile_size;
    tp.target_overlap = overlap;
    auto t0 = std::chrono::steady_clock::now();
    sd::Tensor<float> r = m->vae->decode(m->n_threads, zs, tp, false, false, false, true);
    auto t1 = std::chrono::steady_clock::now();
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (r.empty()) return fail("vae decode failed");
    return from_sd(r, out);
}

int sdh_model_dump_graph(sdh_model* m, const sdh_tensor* x, const sdh_tensor* timesteps, const sdh_tensor* context,
                         const sdh_tensor* y, const char* path) {
    if (!m || !x) return fail("null argument");
    auto xs = to_sd_nd(x, 4);
    auto ts = to_sd_nd(timesteps, 1);
    auto cs = to_sd_nd(context, 3);
    auto ys = to_sd_nd(y, 2);
    ggml_cgraph* gf = build_only(m, xs, ts, cs, ys);
    if (!gf) return fail("graph build failed");
    m->last_flops = graph_flops(gf);
    m->last_nodes = ggml_graph_n_nodes(gf);
    FILE* f = path ? fopen(path, "w") : nullptr;
    if (f) {
        int n = ggml_graph_n_nodes(gf);
        for (int i = 0; i < n; ++i) {
            ggml_tensor* nd = ggml_graph_node(gf, i);
            fprintf(f, "%d %s", i, ggml_op_name(nd->op));
            if (nd->op == GGML_OP_UNARY) fprintf(f, "(%s)", ggml_unary_op_name(ggml_get_unary_op(nd)));
            fprintf(f, " %s [%lld,%lld,%lld,%lld] nb[%zu,%zu,%zu,%zu] fl=%d view=%d p0=%d p1=%d", ggml_type_name(nd->type),
                    (long long)nd->ne[0], (long long)nd->ne[1], (long long)nd->ne[2], (long long)nd->ne[3], nd->nb[0], nd->nb[1],
                    nd->nb[2], nd->nb[3], nd->flags, nd->view_src ? 1 : 0, nd->op_params[0], nd->op_params[1]);
            for (int s = 0; s < GGML_MAX_SRC; ++s) {
                ggml_tensor* sr = nd->src[s];
                if (!sr) continue;
                fprintf(f, " | %s [%lld,%lld,%lld,%lld] nb[%zu,%zu,%zu,%zu] %s%s", ggml_type_name(sr->type), (long long)sr->ne[0],
                        (long long)sr->ne[1], (long long)sr->ne[2], (long long)sr->ne[3], sr->nb[0], sr->nb[1], sr->nb[2], sr->nb[3],
                        ggml_op_name(sr->op), ggml_is_contiguous(sr) ? "" : " NC");
            }
            fprintf(f, "\n");
        }
        fclose(f);
    }
    return m->last_nodes;
}

// In-place update of a model weight THROUGH A GRAPH on the model's backend -- what the reference's LoRA apply does
// (src/lora.hpp:934-937: ggml_add_inplace(model_tensor, updown) computed on the runtime backend): w += value for the first parameter whose
// name contains `name_substr` and has `n_dims` dimensions.  A backend that keeps derived copies of weights must notice the write.
int sdh_model_add_to_weight(sdh_model* m, const char* name_substr, int n_dims, float value) {
    if (!m || !name_substr) return fail("null argument");
    ggml_tensor* w = nullptr;
    for (auto& kv : m->params)
        if (kv.first.find(name_substr) != std::string::npos && ggml_n_dims(kv.second) == n_dims) { w = kv.second; break; }
    if (!w) return fail(std::string("no parameter matches ") + name_substr);
    ggml_init_params ip = {ggml_tensor_overhead() * 8 + ggml_graph_overhead_custom(16, false), nullptr, true};
    ggml_context* ctx = ggml_init(ip);
    if (!ctx) return fail("ggml_init failed");
    ggml_tensor* delta = ggml_new_tensor_4d(ctx, GGML_TYPE_F32, w->ne[0], w->ne[1], w->ne[2], w->ne[3]);
    ggml_tensor* r = ggml_add_inplace(ctx, w, delta);
    ggml_cgraph* gf = ggml_new_graph_custom(ctx, 16, false);
    ggml_build_forward_expand(gf, r);
    int rc = 0;
    if (!ggml_backend_supports_op(m->backend, r)) rc = fail("in-place ADD into this weight type is not supported by the device");
    ggml_backend_buffer_t buf = rc == 0 ? ggml_backend_alloc_ctx_tensors(ctx, m->backend) : nullptr;
    if (rc == 0 && !buf) rc = fail("buffer allocation failed");
    if (buf) {
        std::vector<float> v((size_t)ggml_nelements(delta), value);
        ggml_backend_tensor_set(delta, v.data(), 0, v.size() * 4);
        if (ggml_backend_graph_compute(m->backend, gf) != GGML_STATUS_SUCCESS) rc = fail("graph_compute failed");
        ggml_backend_synchronize(m->backend);
        ggml_backend_buffer_free(buf);
    }
    const int n = (int)std::min<int64_t>(ggml_nelements(w), 0x7fffffff);
    ggml_free(ctx);
    return rc == 0 ? n : rc;
}

// ---------------------------------------------------------------- graph export for the f64 arbiter (oracle/graph_f64.py)
// The graph the reference builds for these inputs, with everything a host-side interpreter needs: per tensor op / type / ne / nb /
// op_params / source ids / view root + offset, and the DATA of every leaf (weights read back from the backend buffer and converted to
// f32 -- F16 / BF16 / Q8_0 values are exactly representable; inputs taken from the runner's pending host copies).
namespace {
struct RunnerPeek : GGMLRunner {
    static std::map<ggml_tensor*, const void*> GGMLRunner::*inputs() { return &RunnerPeek::backend_tensor_data_map; }
};
GGMLRunner* runner_of(sdh_model* m) {
    switch (m->arch) {
        case ARCH_UNET: return m->unet.get();
        case ARCH_VAE: return m->vae.get();
        case ARCH_FLUX: return m->flux.get();
        case ARCH_MMDIT: return m->mmdit.get();
        case ARCH_WAN: return m->wan.get();
        case ARCH_CLIP: return m->clip.get();
        case ARCH_T5: return m->t5.get();
        case ARCH_WANVAE: return m->wan_vae.get();
    }
    return nullptr;
}
}  // namespace

int sdh_model_export_graph(sdh_model* m, const sdh_tensor* x, const sdh_tensor* timesteps, const sdh_tensor* context, const sdh_tensor* y,
                           const char* path_prefix) {
    if (!m || !x || !path_prefix) return fail("null argument");
    auto xs = to_sd_nd(x, 4);
    auto ts = to_sd_nd(timesteps, 1);
    auto cs = to_sd_nd(context, 3);
    auto ys = to_sd_nd(y, 2);
    ggml_cgraph* gf = build_only(m, xs, ts, cs, ys);
    if (!gf) return fail("graph build failed");
    GGMLRunner* runner = runner_of(m);
    const auto& pending = runner->*RunnerPeek::inputs();
    std::vector<ggml_tensor*> order;
    std::map<const ggml_tensor*, int> id;
    std::function<void(ggml_tensor*)> visit = [&](ggml_tensor* t) {
        if (!t || id.count(t)) return;
        if (t->view_src) visit(t->view_src);
        for (int s = 0; s < GGML_MAX_SRC; ++s) visit(t->src[s]);
        id[t] = (int)order.size();
        order.push_back(t);
    };
    const int n = ggml_graph_n_nodes(gf);
    for (int i = 0; i < n; ++i) visit(ggml_graph_node(gf, i));
    std::string jpath = std::string(path_prefix) + ".json", bpath = std::string(path_prefix) + ".bin";
    FILE* fj = fopen(jpath.c_str(), "w");
    FILE* fb = fopen(bpath.c_str(), "wb");
    if (!fj || !fb) { if (fj) fclose(fj); if (fb) fclose(fb); return fail("cannot open export files"); }
    fprintf(fj, "{\"tensors\":[\n");
    size_t off = 0;
    std::vector<float> f32;
    std::vector<uint8_t> raw;
    for (size_t k = 0; k < order.size(); ++k) {
        ggml_tensor* t = order[k];
        long long data_off = -1;
        if (t->op == GGML_OP_NONE && !t->view_src) {
            const int64_t ne = ggml_nelements(t);
            f32.assign((size_t)ne, 0.f);
            auto it = pending.find(t);
            const void* src = nullptr;
            if (it != pending.end()) src = it->second;
            else if (t->buffer) { raw.resize(ggml_nbytes(t)); ggml_backend_tensor_get(t, raw.data(), 0, raw.size()); src = raw.data(); }
            if (src) {
                if (t->type == GGML_TYPE_F32) memcpy(f32.data(), src, (size_t)ne * 4);
                else if (t->type == GGML_TYPE_I32) { const int32_t* p = (const int32_t*)src; for (int64_t i = 0; i < ne; ++i) f32[i] = (float)p[i]; }
                else {
                    const ggml_type_traits* tt = ggml_get_type_traits(t->type);
                    if (!tt->to_float) { fclose(fj); fclose(fb); return fail("leaf type without to_float"); }
                    tt->to_float(src, f32.data(), ne);
                }
                data_off = (long long)off;
                fwrite(f32.data(), 4, (size_t)ne, fb);
                off += (size_t)ne;
            }
        }
        fprintf(fj, "{\"id\":%d,\"op\":\"%s\",\"uop\":\"%s\",\"type\":\"%s\",\"ne\":[%lld,%lld,%lld,%lld],\"nb\":[%zu,%zu,%zu,%zu],\"params\":[", (int)k,
                ggml_op_name(t->op), t->op == GGML_OP_UNARY ? ggml_unary_op_name(ggml_get_unary_op(t)) : "", ggml_type_name(t->type), (long long)t->ne[0],
                (long long)t->ne[1], (long long)t->ne[2], (long long)t->ne[3], t->nb[0], t->nb[1], t->nb[2], t->nb[3]);
        for (int q = 0; q < 16; ++q) fprintf(fj, "%d%s", t->op_params[q], q == 15 ? "" : ",");
        fprintf(fj, "],\"src\":[");
        bool first = true;
        for (int s = 0; s < GGML_MAX_SRC; ++s) {
            if (!t->src[s]) { if (s < 4) { fprintf(fj, "%s-1", first ? "" : ","); first = false; } continue; }
            fprintf(fj, "%s%d", first ? "" : ",", id[t->src[s]]);
            first = false;
        }
        fprintf(fj, "],\"view_src\":%d,\"view_offs\":%zu,\"data\":%lld,\"name\":\"%s\",\"flags\":%d}%s\n", t->view_src ? id[t->view_src] : -1, t->view_offs, data_off,
                t->name, t->flags, k + 1 == order.size() ? "" : ",");
    }
    fprintf(fj, "],\n\"nodes\":[");
    for (int i = 0; i < n; ++i) fprintf(fj, "%d%s", id[ggml_graph_node(gf, i)], i + 1 == n ? "" : ",");
    fprintf(fj, "],\n\"result\":%d}\n", id[ggml_graph_node(gf, n - 1)]);
    fclose(fj);
    fclose(fb);
    return n;
}

double sdh_model_last_graph_flops(const sdh_model* m) { return m->last_flops; }
int sdh_model_last_graph_nodes(const sdh_model* m) { return m->last_nodes; }

// ---------------------------------------------------------------- scheduler + sampler
// calculate_alphas_cumprod / refresh_compvis_denoiser_sigmas live in a .cpp of the reference
// (src/stable-diffusion.cpp:173-186, 666-681), not in a header; these few lines restate them.
static void init_compvis(CompVisDenoiser& d) {
    float ls_sqrt = sqrtf(0.00085f), le_sqrt = sqrtf(0.0120f);
    float amount = le_sqrt - ls_sqrt, product = 1.0f;
    for (int i = 0; i < TIMESTEPS; i++) {
        float beta = ls_sqrt + amount * ((float)i / (TIMESTEPS - 1));
        product *= 1.0f - powf(beta, 2.0f);
        float ac       = product;
        d.sigmas[i]     = std::sqrt((1 - ac) / ac);
        d.log_sigmas[i] = std::log(d.sigmas[i]);
    }
}

int sdh_schedule(int steps, float* sigmas, float* timesteps) {
    CompVisDenoiser d;
    init_compvis(d);
    std::vector<float> s = d.get_sigmas((uint32_t)steps, 0, DISCRETE_SCHEDULER, VERSION_SD1);
    if ((int)s.size() != steps + 1) return fail("scheduler returned unexpected length");
    for (int i = 0; i <= steps; ++i) sigmas[i] = s[i];
    if (timesteps) for (int i = 0; i < steps; ++i) timesteps[i] = d.sigma_to_t(s[i]);
    return 0;
}

int sdh_randn(uint64_t seed, float* dst, size_t n) {
    PhiloxRNG rng;
    rng.manual_seed(seed);
    std::vector<float> v = rng.randn((uint32_t)n);
    memcpy(dst, v.data(), n * sizeof(float));
    return 0;
}

int sdh_sample(sdh_model* m, const char* method_s, int steps, float cfg_scale, float eta, uint64_t sampler_seed,
               const sdh_tensor* noise, const sdh_tensor* cond, const sdh_tensor* uncond, const sdh_tensor* y_cond,
               const sdh_tensor* y_uncond, sdh_tensor* out, float* sigmas_out, float* timesteps_out, int* n_forwards,
               double* wall_ms) {
    return sdh_sample_split(m, method_s, steps, cfg_scale, eta, sampler_seed, noise, cond, uncond, y_cond, y_uncond, out, sigmas_out,
                            timesteps_out, n_forwards, wall_ms, -1, nullptr, nullptr);
}

int sdh_model_backend_stats(sdh_model* m, double* out, int n) {
    if (!m || !out) return fail("null argument");
    ggml_backend_dev_t dev = ggml_backend_get_device(m->backend);
    ggml_backend_reg_t reg = dev ? ggml_backend_dev_backend_reg(dev) : nullptr;
    typedef int (*get_stats_t)(ggml_backend_t, void*);
    get_stats_t fn = reg ? (get_stats_t)ggml_backend_reg_get_proc_address(reg, "ggml_backend_b200_get_stats") : nullptr;
    if (!fn) return fail("backend has no ggml_backend_b200_get_stats");
    struct { uint64_t graphs, launches, nodes, fused; double last_ms, total_ms; uint64_t tc, reserved[8], ext[16], side; } s;
    if (fn(m->backend, &s) != 0) return fail("get_stats failed");
    double v[32] = {(double)s.graphs, (double)s.launches, (double)s.nodes, (double)s.fused, s.last_ms, s.total_ms, (double)s.tc,
                    (double)s.reserved[0], (double)s.reserved[1], (double)s.reserved[2], (double)s.reserved[3], (double)s.reserved[4],
                    (double)s.reserved[5], (double)s.reserved[6], (double)s.reserved[7], 0};
    v[15] = (double)s.side;
    for (int i = 0; i < 16; ++i) v[16 + i] = (double)s.ext[i];
    for (int i = 0; i < n && i < 32; ++i) out[i] = v[i];
    return 0;
}

// CFG-split exchange through the backend's peer mailbox (include/ggml-b200.h ggml_backend_b200_peer_mailbox_*)
static void* b200_proc(sdh_model* m, const char* name) {
    ggml_backend_dev_t dev = ggml_backend_get_device(m->backend);
    ggml_backend_reg_t reg = dev ? ggml_backend_dev_backend_reg(dev) : nullptr;
    return reg ? ggml_backend_reg_get_proc_address(reg, name) : nullptr;
}
int sdh_model_mailbox_create(sdh_model* m, size_t bytes, void* handle_out64) {
    if (!m) return fail("null argument");
    auto fn = (int (*)(ggml_backend_t, size_t, void*))b200_proc(m, "ggml_backend_b200_peer_mailbox_create");
    if (!fn) return fail("backend has no peer mailbox");
    return fn(m->backend, bytes, handle_out64) == 0 ? 0 : fail("mailbox create failed");
}
int sdh_model_mailbox_connect(sdh_model* m, const void* peer_handle64) {
    if (!m) return fail("null argument");
    auto fn = (int (*)(ggml_backend_t, const void*))b200_proc(m, "ggml_backend_b200_peer_mailbox_connect");
    if (!fn) return fail("backend has no peer mailbox");
    if (fn(m->backend, peer_handle64) != 0) return fail("mailbox connect failed");
    m->mailbox = true;
    return 0;
}
void sdh_model_mailbox_close(sdh_model* m) {
    if (!m) return;
    auto fn = (void (*)(ggml_backend_t))b200_proc(m, "ggml_backend_b200_peer_mailbox_close");
    if (fn) fn(m->backend);
    m->mailbox = false;
}

int sdh_model_set_backend_option(sdh_model* m, const char* key, int value) {
    if (!m) return fail("null argument");
    ggml_backend_dev_t dev = ggml_backend_get_device(m->backend);
    ggml_backend_reg_t reg = dev ? ggml_backend_dev_backend_reg(dev) : nullptr;
    typedef int (*set_opt_t)(ggml_backend_t, const char*, int);
    set_opt_t fn = reg ? (set_opt_t)ggml_backend_reg_get_proc_address(reg, "ggml_backend_b200_set_option") : nullptr;
    if (!fn) return fail("backend has no ggml_backend_b200_set_option");
    return fn(m->backend, key, value);
}

int sdh_sample_split(sdh_model* m, const char* method_s, int steps, float cfg_scale, float eta, uint64_t sampler_seed,
                     const sdh_tensor* noise, const sdh_tensor* cond, const sdh_tensor* uncond, const sdh_tensor* y_cond,
                     const sdh_tensor* y_uncond, sdh_tensor* out, float* sigmas_out, float* timesteps_out, int* n_forwards,
                     double* wall_ms, int role, sdh_exchange_fn exchange, void* user) {
    if (!m || m->arch != ARCH_UNET) return fail("sdh_sample needs a unet model");
    auto denoiser = std::make_shared<CompVisDenoiser>();
    init_compvis(*denoiser);
    std::vector<float> sigmas = denoiser->get_sigmas((uint32_t)steps, 0, DISCRETE_SCHEDULER, m->version);
    if (sigmas_out) for (size_t i = 0; i < sigmas.size(); ++i) sigmas_out[i] = sigmas[i];

    sd::Tensor<float> noise_t = to_sd_nd(noise, 4);
    sd::Tensor<float> cond_t = to_sd_nd(cond, 3), uncond_t = to_sd_nd(uncond, 3);
    sd::Tensor<float> yc = to_sd_nd(y_cond, 2), yu = to_sd_nd(y_uncond, 2);
    sd::Tensor<float> init_latent = sd::Tensor<float>::zeros_like(noise_t);
    sd::Tensor<float> x_t = denoiser->noise_scaling(sigmas[0], noise_t, init_latent);  // stable-diffusion.cpp:2620-2622

    auto rng = std::make_shared<PhiloxRNG>();
    rng->manual_seed(sampler_seed);
    sd::guidance::ClassifierFreeGuidance cfg(cfg_scale, 1.0f);
    int forwards = 0;
    bool failed  = false;

    // minimal restatement of the denoise lambda, src/stable-diffusion.cpp:2625-2895 (plain CFG path)
    denoise_cb_t denoise = [&](const sd::Tensor<float>& x, float sigma, int step) -> sd::guidance::GuiderOutput {
        std::vector<float> scaling = denoiser->get_scalings(sigma);
        float c_skip = scaling[0], c_out = scaling[1], c_in = scaling[2];
        float t = denoiser->sigma_to_t(sigma);  // prepare_sample_timesteps, :2411-2434
        if (timesteps_out && step >= 1 && step <= steps) timesteps_out[step - 1] = t;
        sd::Tensor<float> timesteps_tensor({1}, std::vector<float>{t});
        sd::Tensor<float> noised_input = x * c_in;
        sd::Tensor<float> cond_out, uncond_out;
        const bool want_uncond = cfg_scale != 1.0f && !uncond_t.empty();
        if (role == 2 && want_uncond) {
            // batched CFG: cond and uncond as ONE forward with N = 2 -- x [W,H,C,2], context [768,77,2],
            // timesteps [2] (unet.hpp:535-540 takes any N) -- then split; the guidance below is unchanged
            auto stack2 = [](const sd::Tensor<float>& a, const sd::Tensor<float>& b) {
                if (a.empty() || b.empty()) return sd::Tensor<float>();
                std::vector<int64_t> shape = a.shape();
                shape.back() *= 2;
                std::vector<float> data(a.values());
                data.insert(data.end(), b.values().begin(), b.values().end());
                return sd::Tensor<float>(shape, std::move(data));
            };
            sd::Tensor<float> both = run_model(m, stack2(noised_input, noised_input), stack2(timesteps_tensor, timesteps_tensor), stack2(cond_t, uncond_t),
                                               stack2(yc, yu));
            forwards++;
            if (both.empty() || both.numel() != 2 * noised_input.numel()) { failed = true; return {}; }
            const int64_t half = noised_input.numel();
            cond_out   = sd::Tensor<float>(noised_input.shape(), std::vector<float>(both.data(), both.data() + half));
            uncond_out = sd::Tensor<float>(noised_input.shape(), std::vector<float>(both.data() + half, both.data() + 2 * half));
        } else if ((role == 0 || role == 1) && !exchange && m->mailbox) {
            // CFG batch split, device-side exchange: this rank evaluates ONE branch; the backend stores the eps prediction into the
            // peer GPU's mailbox over NVLink at the end of the forward and waits for the peer's (kernels/peer.cu).  The host only reads
            // its own result (as always) and the peer's payload from the local mailbox.
            sd::Tensor<float> mine = role == 0 ? run_model(m, noised_input, timesteps_tensor, cond_t, yc)
                                               : run_model(m, noised_input, timesteps_tensor, uncond_t, yu);
            forwards++;
            if (mine.empty()) { failed = true; return {}; }
            sd::Tensor<float> other = sd::Tensor<float>::zeros_like(mine);
            auto rd = (int (*)(ggml_backend_t, void*))b200_proc(m, "ggml_backend_b200_peer_mailbox_read");
            if (!rd || rd(m->backend, other.data()) != 0) { failed = true; return {}; }
            cond_out   = role == 0 ? mine : other;
            uncond_out = role == 0 ? other : mine;
        } else if (role < 0 || role == 2 || !exchange) {
            cond_out = run_model(m, noised_input, timesteps_tensor, cond_t, yc);  // :2811
            forwards++;
            if (cond_out.empty()) { failed = true; return {}; }
            if (want_uncond) {
                uncond_out = run_model(m, noised_input, timesteps_tensor, uncond_t, yu);  // :2829
                forwards++;
                if (uncond_out.empty()) { failed = true; return {}; }
            }
        } else {
            // CFG batch split over a pair of GPUs: this rank evaluates ONE branch; the caller's
            // exchange callback all-gathers the eps prediction (the "single all-gather on the latent")
            sd::Tensor<float> mine = role == 0 ? run_model(m, noised_input, timesteps_tensor, cond_t, yc)
                                               : run_model(m, noised_input, timesteps_tensor, uncond_t, yu);
            forwards++;
            if (mine.empty()) { failed = true; return {}; }
            cond_out   = sd::Tensor<float>::zeros_like(mine);
            uncond_out = sd::Tensor<float>::zeros_like(mine);
            if (exchange(mine.data(), cond_out.data(), uncond_out.data(), (size_t)mine.numel(), user) != 0) { failed = true; return {}; }
        }
        sd::guidance::GuidanceInput gi;
        gi.step          = step;
        gi.schedule_size = sigmas.size();
        gi.pred_cond     = &cond_out;
        gi.pred_uncond   = uncond_out.empty() ? nullptr : &uncond_out;
        sd::guidance::GuiderOutput guided = cfg.forward(gi, {});
        sd::guidance::GuiderOutput o;
        o.pred = guided.pred * c_out + x * c_skip;  // :2876
        return o;
    };

    std::string ms = method_s ? method_s : "euler_a";
    sample_method_t method = ms == "euler" ? EULER_SAMPLE_METHOD : EULER_A_SAMPLE_METHOD;
    auto t0 = std::chrono::steady_clock::now();
    sd::Tensor<float> x0 = sample_k_diffusion(method, denoise, x_t, sigmas, rng, eta, false, nullptr, denoiser);
    auto t1 = std::chrono::steady_clock::now();
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (n_forwards) *n_forwards = forwards;
    if (failed || x0.empty()) return fail("sampling failed");
    return from_sd(x0, out);
}

}  // extern "C"

// build_graph entry points differ per runner; all are public in the reference
int sdh_model_check_ops(sdh_model* m, const sdh_tensor* x, const sdh_tensor* timesteps, const sdh_tensor* context, const sdh_tensor* y,
                        sdh_op_supported_fn fn, char* first_unsupported, size_t len) {
    if (!m || !x || !fn) return fail("null argument");
    auto xs = to_sd_nd(x, 4);
    auto ts = to_sd_nd(timesteps, 1);
    auto cs = to_sd_nd(context, 3);
    auto ys = to_sd_nd(y, 2);
    ggml_cgraph* gf = build_only(m, xs, ts, cs, ys);
    if (!gf) return fail("graph build failed");
    int bad = 0;
    for (int i = 0; i < ggml_graph_n_nodes(gf); ++i) {
        ggml_tensor* nd = ggml_graph_node(gf, i);
        if (fn(nd)) continue;
        if (bad == 0 && first_unsupported && len > 0)
            snprintf(first_unsupported, len, "node %d %s %s [%lld,%lld,%lld,%lld] src0 %s", i, ggml_op_name(nd->op), ggml_type_name(nd->type), (long long)nd->ne[0],
                     (long long)nd->ne[1], (long long)nd->ne[2], (long long)nd->ne[3], nd->src[0] ? ggml_type_name(nd->src[0]->type) : "-");
        bad++;
    }
    m->last_nodes = ggml_graph_n_nodes(gf);
    return bad;
}

static ggml_cgraph* build_only(sdh_model* m, const sd::Tensor<float>& x, const sd::Tensor<float>& t,
                               const sd::Tensor<float>& ctx, const sd::Tensor<float>& y) {
    switch (m->arch) {
        case ARCH_UNET:
            m->unet->reset_compute_ctx();
            return m->unet->build_graph(x, t, ctx, {}, y);
        case ARCH_VAE:
            m->vae->reset_compute_ctx();
            return m->vae->build_graph(x, true);
        case ARCH_FLUX:
            m->flux->reset_compute_ctx();
            return m->flux->build_graph(x, t, ctx, {}, y);
        case ARCH_MMDIT:
            m->mmdit->reset_compute_ctx();
            return m->mmdit->build_graph(x, t, ctx, y);
        case ARCH_WAN:
            m->wan->reset_compute_ctx();
            return m->wan->build_graph(x, t, ctx);
        case ARCH_CLIP:
            m->clip->reset_compute_ctx();
            return m->clip->build_graph(token_ids(x));
        case ARCH_T5:
            m->t5->reset_compute_ctx();
            return m->t5->build_graph(token_ids(x));
        case ARCH_WANVAE:
            m->wan_vae->reset_compute_ctx();
            return m->wan_vae->build_graph(x, true);
    }
    return nullptr;
}

// ---------------------------------------------------------------- single-op runner (parity tests)
extern "C" int sdh_run_op(const char* device, const char* op_s, int n_in, const sdh_tensor* in, const int32_t* itypes, const int32_t* ip,
                          const float* fp, sdh_tensor* out, int n_threads) {
    std::string op = op_s ? op_s : "";
    ggml_backend_t be = init_device(device ? device : "CPU", n_threads);
    if (!be) return fail(std::string("no such device: ") + (device ? device : "null"));
    ggml_init_params ip0 = {ggml_tensor_overhead() * 256 + ggml_graph_overhead(), nullptr, true};
    ggml_context* ctx = ggml_init(ip0);
    ggml_tensor* t[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < n_in && i < 4; ++i) {
        if (!in[i].data) continue;
        t[i] = ggml_new_tensor_4d(ctx, (ggml_type)(itypes ? itypes[i] : 0), in[i].ne[0], in[i].ne[1], in[i].ne[2], in[i].ne[3]);
        ggml_set_input(t[i]);
    }
    ggml_tensor* r = nullptr;
    auto I = [&](int k) { return ip ? ip[k] : 0; };
    auto F = [&](int k) { return fp ? fp[k] : 0.f; };
    if (op == "mul_mat") r = ggml_mul_mat(ctx, t[0], t[1]);
    else if (op == "conv_2d") {
        r = ggml_conv_2d(ctx, t[0], t[1], I(0), I(1), I(2), I(3), I(4), I(5));
        if (t[2]) r = ggml_add_inplace(ctx, r, t[2]);
    } else if (op == "rope") {
        // Rope::apply_rope (rope.hpp:966-1010) on x [d, H, L, N] with pe [2, 2, d/2, L]; optional QKNorm in front (RMSNorm(eps = fp[0]) * w = t[2],
        // flux.hpp:213-261) and the K-side cast to F16 behind (ip[0])
        ggml_tensor* xx = t[0];
        if (t[2]) xx = ggml_mul(ctx, ggml_rms_norm(ctx, xx, F(0)), t[2]);
        r = Rope::apply_rope(ctx, xx, t[1], true);
        if (I(0)) r = ggml_cast(ctx, r, GGML_TYPE_F16);
    } else if (op == "conv_3d") {
        // w [OC*IC, KD, KH, KW], x [N*IC, ID, IH, IW]; ip = {IC, s0,s1,s2, p0,p1,p2, d0,d1,d2} (Wan patch embedding / causal 3-D convs)
        r = ggml_conv_3d(ctx, t[0], t[1], I(0), I(1), I(2), I(3), I(4), I(5), I(6), I(7), I(8), I(9));
    } else if (op == "gn_silu_conv") {
        // ResBlock prologue + conv (block.hpp:124-150): GroupNorm32 -> SiLU -> Conv2d 3x3 (x, gn_w, gn_b, conv_w)
        ggml_tensor* hh = ggml_group_norm(ctx, t[0], I(0), F(0));
        hh = ggml_mul_inplace(ctx, hh, t[1]);
        hh = ggml_add_inplace(ctx, hh, t[2]);
        if (I(1)) hh = ggml_silu_inplace(ctx, hh);
        r = ggml_conv_2d(ctx, t[3], hh, 1, 1, I(2), I(2), 1, 1);
    } else if (op == "upscale_conv") {
        // UpSampleBlock (block.hpp:58-65): nearest x2 -> Conv2d 3x3 (+ bias)   (x, conv_w, bias)
        ggml_tensor* hh = ggml_upscale(ctx, t[0], 2, GGML_SCALE_MODE_NEAREST);
        r = ggml_conv_2d(ctx, t[1], hh, 1, 1, 1, 1, 1, 1);
        if (t[2]) r = ggml_add_inplace(ctx, r, t[2]);
    } else if (op == "im2col") r = ggml_im2col(ctx, t[0], t[1], I(0), I(1), I(2), I(3), I(4), I(5), true, (ggml_type)I(6));
    else if (op == "group_norm") {
        r = ggml_group_norm(ctx, t[0], I(0), F(0));
        if (t[1]) r = ggml_mul_inplace(ctx, r, t[1]);
        if (t[2]) r = ggml_add_inplace(ctx, r, t[2]);
        if (I(1)) r = ggml_silu_inplace(ctx, r);
    } else if (op == "norm") r = ggml_norm(ctx, t[0], F(0));
    else if (op == "rms_norm") r = ggml_rms_norm(ctx, t[0], F(0));
    else if (op == "soft_max") r = ggml_soft_max_ext(ctx, t[0], t[1], F(0), F(1));
    else if (op == "flash_attn") {
        r = ggml_flash_attn_ext(ctx, t[0], t[1], t[2], t[3], F(0), 0.f, 0.f);
        ggml_flash_attn_ext_set_prec(r, GGML_PREC_F32);
    } else if (op == "attention") r = ggml_ext_attention_ext(ctx, be, t[0], t[1], t[2], I(0), nullptr, false, I(1) != 0);
    else if (op == "upscale") r = ggml_upscale(ctx, t[0], I(0), (ggml_scale_mode)I(1));
    else if (op == "timestep_embedding") r = ggml_timestep_embedding(ctx, t[0], I(0), I(1));
    else if (op == "unary") r = ggml_unary(ctx, t[0], (ggml_unary_op)I(0));
    else if (op == "add") r = ggml_add(ctx, t[0], t[1]);
    else if (op == "mul") r = ggml_mul(ctx, t[0], t[1]);
    else if (op == "scale") r = ggml_scale_bias(ctx, t[0], F(0), F(1));
    else if (op == "concat") r = ggml_concat(ctx, t[0], t[1], I(0));
    else if (op == "cont_permute") r = ggml_cont(ctx, ggml_permute(ctx, t[0], I(0), I(1), I(2), I(3)));
    else if (op == "cpy") r = ggml_cast(ctx, t[0], (ggml_type)I(0));
    if (!r) { ggml_free(ctx); ggml_backend_free(be); return fail("unknown op: " + op); }
    if (r->type != GGML_TYPE_F32) r = ggml_cast(ctx, r, GGML_TYPE_F32);
    if (!ggml_is_contiguous(r)) r = ggml_cont(ctx, r);
    ggml_set_output(r);
    for (int i = 0; i < 4; ++i) out->ne[i] = r->ne[i];
    int rc = 0;
    if (out->data) {
        ggml_cgraph* gf = ggml_new_graph_custom(ctx, 64, false);
        ggml_build_forward_expand(gf, r);
        bool supported = true;
        for (int i = 0; i < ggml_graph_n_nodes(gf); ++i)
            if (!ggml_backend_supports_op(be, ggml_graph_node(gf, i))) {
                supported = false;
                rc = fail(std::string("op not supported by device: ") + ggml_op_name(ggml_graph_node(gf, i)->op));
            }
        ggml_backend_buffer_t buf = supported ? ggml_backend_alloc_ctx_tensors(ctx, be) : nullptr;
        if (supported && !buf) rc = fail("buffer allocation failed");
        if (buf) {
            std::vector<uint8_t> conv;
            for (int i = 0; i < 4; ++i) {
                if (!t[i]) continue;
                int64_t n = ggml_nelements(t[i]);
                if (t[i]->type == GGML_TYPE_F32) ggml_backend_tensor_set(t[i], in[i].data, 0, n * 4);
                else {
                    conv.resize(ggml_nbytes(t[i]));
                    const ggml_type_traits* tt = ggml_get_type_traits(t[i]->type);
                    int64_t nrows = n / t[i]->ne[0];
                    size_t rb = ggml_row_size(t[i]->type, t[i]->ne[0]);
                    for (int64_t rr = 0; rr < nrows; ++rr) tt->from_float_ref(in[i].data + rr * t[i]->ne[0], conv.data() + rr * rb, t[i]->ne[0]);
                    ggml_backend_tensor_set(t[i], conv.data(), 0, conv.size());
                }
            }
            if (ggml_backend_graph_compute(be, gf) != GGML_STATUS_SUCCESS) rc = fail("graph_compute failed");
            else ggml_backend_tensor_get(r, out->data, 0, ggml_nbytes(r));
            ggml_backend_buffer_free(buf);
        }
    }
    ggml_free(ctx);
    ggml_backend_free(be);
    return rc;
}
