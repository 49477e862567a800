// gemm_bench.cu -- micro-benchmark of the tcgen05 GEMM / implicit-conv kernels at the SD1.5 UNet's shapes.
// Steady state (warm L2, back-to-back launches on one stream, CUDA events), prints us / TFLOP/s per shape.
// Tuning aid only: not part of the product path.  usage: gemm_bench [reps]
#include "../csrc/b200_ops.h"
#include <cuda_fp16.h>
#include <vector>
#include <string>
#include <cstring>
#include <cmath>

struct Shape { const char* name; int64_t M, N, K, batch; };

int main(int argc, char** argv) {
    int reps = argc > 1 ? atoi(argv[1]) : 50;
    setenv("GGML_B200_GEMM2", "0", 1);      // the one-CTA kernel is the reference of the comparisons below
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    b200_device_info dev{};
    dev.id = 0; dev.sm_count = p.multiProcessorCount;
    std::vector<Shape> shapes = {
        {"conv 64x64 320->320 (im2col)", 4096, 320, 2880, 1}, {"conv 32x32 640->640", 1024, 640, 5760, 1}, {"conv 16x16 1280->1280", 256, 1280, 11520, 1},
        {"conv 8x8 1280->1280", 64, 1280, 11520, 1}, {"conv 32x32 1920->640", 1024, 640, 17280, 1}, {"conv 64x64 960->320", 4096, 320, 8640, 1},
        {"linear qkv L4096 C320", 320, 4096, 320, 1}, {"linear geglu L4096", 2560, 4096, 320, 1}, {"linear ff-out L4096", 320, 4096, 1280, 1},
        {"linear qkv L1024 C640", 640, 1024, 640, 1}, {"linear geglu L1024", 5120, 1024, 640, 1}, {"linear ff-out L1024", 640, 1024, 2560, 1},
        {"linear qkv L256 C1280", 1280, 256, 1280, 1}, {"linear geglu L256", 10240, 256, 1280, 1}, {"linear ff-out L256", 1280, 256, 5120, 1},
        {"linear kv ctx77 C1280", 1280, 77, 768, 1}, {"linear kv ctx77 C320", 320, 77, 768, 1}, {"time emb", 1280, 1, 320, 1}, {"big square", 4096, 4096, 4096, 1},
        {"flux qkv 4352x9216x3072", 9216, 4352, 3072, 1}, {"flux linear2 4352x3072x15360", 3072, 4352, 15360, 1}, {"vae conv 512^2 128->128 (as gemm)", 262144, 128, 1152, 1},
        {"vae conv 256^2 256->256 (as gemm)", 65536, 256, 2304, 1}, {"sd15 conv 64x64 320 x2 (M 8192)", 8192, 320, 2880, 1},
    };
    size_t maxA = 0, maxB = 0, maxD = 0;
    for (auto& s : shapes) { maxA = std::max(maxA, (size_t)s.M * s.K); maxB = std::max(maxB, (size_t)s.N * s.K); maxD = std::max(maxD, (size_t)s.M * s.N); }
    __half *A, *B; float *D, *bias;
    cudaMalloc(&A, maxA * 2); cudaMalloc(&B, maxB * 2); cudaMalloc(&D, maxD * 4); cudaMalloc(&bias, 1 << 20);
    {   // deterministic non-trivial operands (a mismatch count against all-zero data would prove nothing)
        std::vector<__half> h(std::max(maxA, maxB));
        uint32_t x = 12345u;
        for (auto& v : h) { x = x * 1664525u + 1013904223u; v = __float2half(((int)(x >> 20) % 17 - 8) / 64.0f); }
        cudaMemcpy(A, h.data(), maxA * 2, cudaMemcpyHostToDevice);
        for (auto& v : h) { x = x * 1664525u + 1013904223u; v = __float2half(((int)(x >> 20) % 13 - 6) / 32.0f); }
        cudaMemcpy(B, h.data(), maxB * 2, cudaMemcpyHostToDevice);
        std::vector<float> hb(1 << 18);
        for (auto& v : hb) { x = x * 1664525u + 1013904223u; v = ((int)(x >> 20) % 9 - 4) / 8.0f; }
        cudaMemcpy(bias, hb.data(), 1 << 20, cudaMemcpyHostToDevice);
    }
    // GEMM_BENCH_COLD=1: the WEIGHT operand of every launch comes from a different copy in a 1 GB pool, so it is read from HBM like the
    // weights of a real forward (1.7 GB per SD1.5 step, each byte once); activations stay warm.  GEMM_BENCH_PF=1: with the up-front L2
    // request of each CTA's weight slab (b200_gemm_args::wprefetch).
    const bool cold = getenv("GEMM_BENCH_COLD") != nullptr;
    const bool pf = getenv("GEMM_BENCH_PF") != nullptr;
    const size_t pool_bytes = (size_t)1 << 30;
    __half* pool = nullptr;
    if (cold) {
        if (cudaMalloc(&pool, pool_bytes) != cudaSuccess) { printf("no memory for the cold pool\n"); return 1; }
        const size_t chunk = std::max(maxA, maxB) * 2;
        for (size_t off = 0; off < pool_bytes; off += chunk) cudaMemcpy((char*)pool + off, A, std::min(chunk, pool_bytes - off), cudaMemcpyDeviceToDevice);
    }
    unsigned long long* trace; cudaMalloc(&trace, 64);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // device time of `reps` back-to-back launches replayed as ONE CUDA graph (how the backend runs them): no host launch cost in the number
    auto time_graph = [&](auto&& launch) -> double {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        launch();                                   // first launch outside the capture (function attributes, tensor-map cache)
        cudaStreamSynchronize(st);
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return -1.0;
        for (int i = 0; i < reps; ++i) launch();
        if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !graph) { cudaGetLastError(); return -1.0; }
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); cudaGetLastError(); return -1.0; }
        cudaGraphLaunch(exec, st);
        cudaStreamSynchronize(st);
        float ms = 0;
        cudaEventRecord(e0, st);
        cudaGraphLaunch(exec, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        return ms * 1e3 / reps;
    };
    const char* only = getenv("GEMM_BENCH_ONLY");      // substring filter on the shape name
    // GEMM_BENCH_CONV=1: the 3x3 convolutions of the UNet / VAE on the CTA-pair kernel, per-tap boxes against halo reuse (3 / 9 taps per
    // ring stage), over tile widths and split-K factors; the halo result is compared with the per-tap one (same products, other order)
    if (getenv("GEMM_BENCH_CONV")) {
        struct Conv { const char* name; int64_t N, H, W, C, OC; };
        const Conv convs[] = {{"sd15 64x64 320->320 x2", 2, 64, 64, 320, 320}, {"sd15 32x32 640->640 x2", 2, 32, 32, 640, 640}, {"sd15 16x16 1280->1280 x2", 2, 16, 16, 1280, 1280},
                              {"sd15 64x64 640->320 x2", 2, 64, 64, 640, 320}, {"sd15 32x32 1280->640 x2", 2, 32, 32, 1280, 640},
                              {"vae 512x512 128->128", 1, 512, 512, 128, 128}, {"vae 256x256 256->256", 1, 256, 256, 256, 256}, {"vae 128x128 512->512", 1, 128, 128, 512, 512},
                              {"sdxl 128x128 320->320", 1, 128, 128, 320, 320}, {"sdxl 64x64 640->640", 1, 64, 64, 640, 640}};
        printf("%-28s %4s %6s %5s | %9s %9s %s\n", "conv", "bn", "splits", "taps", "us", "TFLOP/s", "max |halo - per-tap|");
        for (const Conv& cv : convs) {
            if (only && !strstr(cv.name, only)) continue;
            const int64_t M = cv.H * cv.W, K = 9 * cv.C;
            if ((size_t)cv.N * M * cv.C > maxA || (size_t)cv.OC * K > maxB || (size_t)cv.N * M * cv.OC > maxD) { printf("%s: buffers too small\n", cv.name); continue; }
            b200_conv_args c; memset(&c, 0, sizeof(c));
            c.x_nhwc = A; c.w_packed = B; c.N = cv.N; c.H = cv.H; c.W = cv.W; c.C = cv.C; c.OC = cv.OC; c.KH = 3; c.KW = 3; c.pad = 1; c.dil = 1; c.D = D; c.bias = bias;
            const double flop = 2.0 * cv.N * M * cv.OC * K;
            std::vector<float> ref, got((size_t)cv.N * M * cv.OC);
            {   // the dispatcher's own choice (what the backend runs)
                const double t = time_graph([&] { b200_launch_conv_tc(st, dev, c, nullptr, 0); });
                printf("%-28s %4s %6s %5s | %9.2f %9.1f   <- dispatcher\n", cv.name, "-", "-", "-", t, flop / t * 1e-6);
            }
            const int bns[] = {256, 192, 160, 128, 96, 64};
            for (int bn : bns) {
                if (bn > 64 && cv.OC <= bn / 2) continue;
                for (int sp = 1; sp <= 4; sp *= 2) {
                    const int64_t tiles = ((M + 255) / 256) * ((cv.OC + bn - 1) / bn) * cv.N;
                    if (sp > 1 && tiles * 2 * sp > 148) continue;
                    const int auto_taps = b200_conv_tc2_halo_taps(bn, sp);
                    const int modes[3] = {0, auto_taps, auto_taps == 9 ? 3 : -1};
                    bool have_ref = false;
                    for (int mi = 0; mi < 3; ++mi) {
                        const int taps = modes[mi];
                        if (taps < 0 || (mi > 0 && taps == 0)) continue;
                        cudaMemsetAsync(D, 0xff, got.size() * 4, st);
                        if (b200_launch_conv_tc2(st, dev, c, bn, sp, taps) <= 0) continue;
                        if (cudaStreamSynchronize(st) != cudaSuccess) { printf("   bn %d splits %d taps %d: FAILED %s\n", bn, sp, taps, cudaGetErrorString(cudaGetLastError())); return 1; }
                        cudaMemcpy(got.data(), D, got.size() * 4, cudaMemcpyDeviceToHost);
                        double maxd = 0;
                        if (taps == 0) { ref = got; have_ref = true; }
                        else if (have_ref) for (size_t i = 0; i < got.size(); ++i) { const double d = fabs((double)got[i] - ref[i]); if (!(d <= maxd)) maxd = d; }
                        const double t = time_graph([&] { b200_launch_conv_tc2(st, dev, c, bn, sp, taps); });
                        printf("%-28s %4d %6d %5d | %9.2f %9.1f   %.3g\n", cv.name, bn, sp, taps, t, flop / t * 1e-6, taps ? maxd : 0.0);
                        fflush(stdout);
                    }
                }
            }
        }
        printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    printf("%-34s %8s %8s %8s | %9s %9s\n", "shape", "M", "N", "K", "us", "TFLOP/s");
    for (auto& s : shapes) {
        if (only) {                                    // comma-separated substrings
            bool hit = false;
            std::string list(only);
            size_t pos = 0;
            while (pos <= list.size()) {
                const size_t c = list.find(',', pos);
                const std::string tok = list.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
                if (!tok.empty() && strstr(s.name, tok.c_str())) hit = true;
                if (c == std::string::npos) break;
                pos = c + 1;
            }
            if (!hit) continue;
        }
        b200_gemm_args g; memset(&g, 0, sizeof(g));
        g.A = A; g.B = B; g.type = GGML_TYPE_F16; g.M = s.M; g.N = s.N; g.K = s.K; g.lda = s.K; g.ldb = s.K; g.batch = 1; g.a_bcast = 1;
        g.a_batch_stride = s.M * s.K; g.b_batch_stride = s.N * s.K; g.d_batch_stride = s.M * s.N; g.D = D; g.ldd = s.M; g.bias = bias; g.bias_mode = 1;
        for (int i = 0; i < 3; ++i) b200_launch_gemm_tc(st, dev, g, nullptr, 0);
        cudaStreamSynchronize(st);
        float ms = 0;
        double us = time_graph([&] { b200_launch_gemm_tc(st, dev, g, nullptr, 0); });
        printf("%-34s %8lld %8lld %8lld | %9.2f %9.1f", s.name, (long long)s.M, (long long)s.N, (long long)s.K, us, 2.0 * s.M * s.N * s.K / us * 1e-6);
        if (cold) {
            // which operand is the weight matrix: the filter (B) of a conv, the [features][K] matrix (A) of a Linear
            const bool w_is_b = strstr(s.name, "conv") != nullptr;
            const size_t wbytes = (size_t)(w_is_b ? s.N : s.M) * s.K * 2;
            const size_t stride = (wbytes + 4095) & ~(size_t)4095;
            const int ncopies = (int)std::min<size_t>(pool_bytes / stride, 64);
            {
                for (int with_pf = 0; with_pf <= (pf ? 1 : 0); ++with_pf) {
                    int it = 0;
                    b200_gemm_args gc = g;
                    gc.wprefetch = with_pf ? (w_is_b ? 2 : 1) : 0;
                    const double t = time_graph([&] {
                        const char* w = (const char*)pool + (size_t)(it++ % ncopies) * stride;
                        if (w_is_b) gc.B = w; else gc.A = w;
                        b200_launch_gemm_tc(st, dev, gc, nullptr, 0);
                    });
                    printf("\n   one-CTA cold weights%s (%d copies of %.1f MB): %8.2f us %7.1f TFLOP/s  weights at %.0f GB/s", with_pf ? " + L2 slab prefetch" : "", ncopies,
                           wbytes / 1e6, t, 2.0 * s.M * s.N * s.K / t * 1e-6, wbytes / t * 1e-3);
                }
            }
            // the pair kernel at the plan the dispatcher's model picks for this shape
            {
                const int nkb = (int)((s.K + 63) / 64);
                double bestc = 1e30; int bbn = 0, bsp = 1;
                const int bns[] = {256, 224, 192, 160, 128, 96, 64, 48, 32};
                for (int bn : bns) {
                    if (bn > 32 && s.N <= bn / 2) continue;
                    const int64_t tiles = ((s.M + 255) / 256) * ((s.N + bn - 1) / bn);
                    for (int sp = 1; sp <= 4; sp *= 2) {
                        if (sp > 1 && (tiles * 2 * sp > 148 || nkb / sp < 4)) break;
                        const double c = b200_gemm_tc2_model(dev, s.M, s.N, 1, nkb, bn, sp);
                        if (c < bestc) { bestc = c; bbn = bn; bsp = sp; }
                    }
                }
                if (bbn && s.M > 128) {
                    for (int with_pf = 0; with_pf <= (pf ? 1 : 0); ++with_pf) {
                        int it = 0;
                        b200_gemm_args gc = g;
                        gc.wprefetch = with_pf ? (w_is_b ? 2 : 1) : 0;
                        const double tw = with_pf ? 0.0 : time_graph([&] { b200_launch_gemm_tc2(st, dev, g, bbn, bsp); });
                        const double t = time_graph([&] {
                            const char* w = (const char*)pool + (size_t)(it++ % ncopies) * stride;
                            if (w_is_b) gc.B = w; else gc.A = w;
                            b200_launch_gemm_tc2(st, dev, gc, bbn, bsp);
                        });
                        if (!with_pf) printf("\n   pair bn %d splits %d warm: %8.2f us", bbn, bsp, tw);
                        printf("\n   pair bn %d splits %d cold weights%s: %8.2f us %7.1f TFLOP/s  weights at %.0f GB/s", bbn, bsp, with_pf ? " + L2 slab prefetch" : "", t,
                               2.0 * s.M * s.N * s.K / t * 1e-6, wbytes / t * 1e-3);
                    }
                }
            }
        }
        // CTA-pair kernel (gemm_tc2.cu) on the same problem: a sweep over (tile N, split-K), each checked element by element against the
        // one-CTA kernel's result
        if (getenv("GEMM_BENCH_PAIR")) {
            std::vector<float> ref((size_t)s.M * s.N), got((size_t)s.M * s.N);
            cudaMemcpy(ref.data(), D, ref.size() * 4, cudaMemcpyDeviceToHost);
            const int bns_all[] = {256, 192, 160, 128, 96, 64};
            const int bns_few[] = {256, 128, 64};
            const bool few = getenv("GEMM_BENCH_FEW") != nullptr;
            std::vector<int> bns(few ? std::begin(bns_few) : std::begin(bns_all), few ? std::end(bns_few) : std::end(bns_all));
            printf("\n");
            for (int bn : bns) {
                if (bn > 64 && s.N <= bn / 2) continue;
                for (int sp = 1; sp <= 4; sp *= 2) {
                    const int64_t tiles = ((s.M + 255) / 256) * ((s.N + bn - 1) / bn);
                    if (sp > 1 && tiles * 2 * sp > 148) continue;
                    cudaMemsetAsync(D, 0xff, ref.size() * 4, st);
                    int ok = b200_launch_gemm_tc2(st, dev, g, bn, sp);
                    if (ok <= 0) continue;
                    if (cudaStreamSynchronize(st) != cudaSuccess) { printf("   pair bn %d splits %d: FAILED %s\n", bn, sp, cudaGetErrorString(cudaGetLastError())); return 1; }
                    cudaMemcpy(got.data(), D, got.size() * 4, cudaMemcpyDeviceToHost);
                    size_t bad = 0; double maxd = 0;
                    for (size_t i = 0; i < ref.size(); ++i) { if (ref[i] != got[i]) { ++bad; double d = fabs((double)ref[i] - got[i]); if (!(d <= maxd)) maxd = d; } }
                    const double t2 = time_graph([&] { b200_launch_gemm_tc2(st, dev, g, bn, sp); });
                    const int nkb = (int)((s.K + 63) / 64);
                    printf("   pair bn %3d splits %d: %8.2f us %7.1f TFLOP/s  model %7.2f us  mismatches %zu (max abs %.3g)\n", bn, sp, t2,
                           2.0 * s.M * s.N * s.K / t2 * 1e-6, b200_gemm_tc2_model(dev, s.M, s.N, 1, nkb, bn, sp) / 1965.0, bad, maxd);
                    fflush(stdout);
                }
            }
        }
        // phase timestamps of CTA (0,0,0): start, setup done, first k-block landed, last k-block landed, accumulator ready, epilogue done, all warps joined, tmem freed
        cudaMemset(trace, 0, 64);
        g.trace = trace;
        cudaEventRecord(e0, st);
        b200_launch_gemm_tc(st, dev, g, nullptr, 0);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long t[8];
        cudaMemcpy(t, trace, 64, cudaMemcpyDeviceToHost);
        printf(" | single %6.2f us; cta0 ns: setup %llu, first-data %llu, mainloop %llu, acc-ready %llu, epilogue %llu, join %llu, free %llu\n", ms * 1e3,
               t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6]);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
