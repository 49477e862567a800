// conv_prep.cu -- producers of the operands the implicit-GEMM convolution (gemm_tc.cu, conv mode) reads through TMA:
//
//   k_gn_stats        per (image, group) mean / rstd of a GroupNorm input (oracle arithmetic: ops.cpp:4079-4152)
//   k_to_nhwc_f16     NCHW f32 activation -> NHWC f16 "shadow" image with the whole ResBlock prologue folded in:
//                     (x - mean_g) * rstd_g * w_c + b_c, SiLU, and optional nearest x2 upsampling
//                     (reference graph: GROUP_NORM, MUL, ADD, SILU[, UPSCALE], IM2COL -- ggml_extend.hpp:1502-1520,
//                     block.hpp:58-65,142).  Rounding to f16 happens exactly where the oracle rounds (im2col output is F16).
//   k_pack_conv_weight  [KW,KH,IC,OC] f16 (ggml) -> [OC][KH][KW][IC] f16 so that K runs (tap, channel) like the NHWC image
//
// All three are HBM/L2 bound: the activation is read once (twice from L2 when statistics are needed) and written once
// at half width; the 0.7 GB/forward of materialised im2col columns of the unfused graph disappear.
#include "../b200_ops.h"
#include "b200_launch.cuh"

#include <cuda_fp16.h>
#include <cstdlib>

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (w == 0) r = warp_sum(r);
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    return red[0];
}

__global__ void __launch_bounds__(1024) k_gn_stats(const float* __restrict__ x, float2* __restrict__ stats, int64_t inner, int C, int cpg, int G,
                                                   float eps, const float* __restrict__ addv) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[32];
    const int g = blockIdx.x, n = blockIdx.y;
    const int c0 = g * cpg, c1 = min(c0 + cpg, C);
    if (c0 >= c1) return;
    const int64_t len = (int64_t)(c1 - c0) * inner;
    const float* xp = x + ((int64_t)n * C + c0) * inner;
    const float* av = addv ? addv + (int64_t)n * C + c0 : nullptr;      // per-channel value added on the fly (x + emb broadcast)
    // (the folded per-channel add keeps the summation ORDER of the plain path -- pairs inside a float4, then the running sum -- so that
    //  folding the ResBlock's broadcast ADD changes no bit: tests/test_gpu_models.py::test_producer_side_fusions_are_bit_identical)
    const bool vec = ((uintptr_t)xp % 16 == 0) && (len % 4 == 0) && (!av || (inner % 4 == 0));
    float s = 0.f;
    if (vec) {
        const float4* x4 = (const float4*)xp;
        for (int64_t i = threadIdx.x; i < len / 4; i += blockDim.x) {
            float4 v = x4[i];
            if (av) { const float e = av[(4 * i) / inner]; v.x += e; v.y += e; v.z += e; v.w += e; }
            s += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) s += av ? xp[i] + av[i / inner] : xp[i];
    }
    const float mean = block_sum(s, red) / (float)len;
    float s2 = 0.f;
    if (vec) {
        const float4* x4 = (const float4*)xp;
        for (int64_t i = threadIdx.x; i < len / 4; i += blockDim.x) {
            float4 v = x4[i];
            if (av) { const float e = av[(4 * i) / inner]; v.x += e; v.y += e; v.z += e; v.w += e; }
            float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
            s2 += (a * a + b * b) + (c * c + d * d);
        }
    } else {
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) { float v = (av ? xp[i] + av[i / inner] : xp[i]) - mean; s2 += v * v; }
    }
    const float var = block_sum(s2, red) / (float)len;
    if (threadIdx.x == 0) stats[(int64_t)n * G + g] = make_float2(mean, 1.0f / sqrtf(var + eps));
}

// Chunked statistics: grid (S, G, N).  A group of the VAE decoder is 1 M floats (512 x 512 x 4 channels); one CTA per group (above)
// leaves 116 of 148 SMs idle and reads the group twice.  Here S CTAs share a group: each keeps its chunk (<= 32 K floats) in
// REGISTERS, computes the chunk mean and the chunk's sum of squared deviations from THAT mean (the oracle's two-pass arithmetic,
// ops.cpp:4079-4152, per chunk), and the last CTA to finish merges the S partial (mean, M2) pairs in chunk order with the exact
// pairwise update (Chan et al.) in double -- deterministic, one read of the activation.
constexpr int GN2_THREADS = 256;                            // ~100 registers per thread: two CTAs per SM
constexpr int GN2_VEC = 16;                                  // float4 per thread
constexpr int GN2_CHUNK = GN2_THREADS * GN2_VEC * 4;         // 16384 floats (64 KB in flight per CTA)
// (32 KB chunks with four CTAs per SM were tried on hardware: the 512 x 512 x 128 VAE level went from 55 us to 72 us -- twice the partial
//  merges and atomics; profiles/r02_summary.md)

__device__ __forceinline__ double block_sum_d(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = (threadIdx.x < (GN2_THREADS >> 5)) ? red[threadIdx.x] : 0.0;
    if (w == 0) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    return red[0];
}

__global__ void __launch_bounds__(GN2_THREADS, 2) k_gn_stats_chunked(const float* __restrict__ x, float2* __restrict__ stats, double2* partial,
                                                                  unsigned* __restrict__ counters, int64_t inner, int C, int cpg, int G, int S,
                                                                  int64_t chunk, float eps, const float* __restrict__ addv) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ double red[32];
    __shared__ int is_last;
    const int ci = blockIdx.x, g = blockIdx.y, n = blockIdx.z;
    const int c0 = g * cpg, c1 = min(c0 + cpg, C);
    const int64_t len = (int64_t)(c1 - c0) * inner;
    const float4* x4 = (const float4*)(x + ((int64_t)n * C + c0) * inner);
    const int64_t e0 = (int64_t)ci * chunk, e1 = min(e0 + chunk, len);        // multiples of 4 (chunk % 4 == 0, len % 4 == 0)
    const int64_t v0 = e0 >> 2, nv = (e1 - e0) >> 2;
    float4 r[GN2_VEC];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < GN2_VEC; ++k) {
        const int64_t i = (int64_t)k * GN2_THREADS + threadIdx.x;
        r[k] = i < nv ? x4[v0 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (addv && i < nv) {       // x + emb broadcast (per channel): inner % 4 == 0, so the four values share a channel
            const float e = addv[(int64_t)n * C + c0 + (4 * (v0 + i)) / inner];
            r[k].x += e; r[k].y += e; r[k].z += e; r[k].w += e;
        }
        s += (r[k].x + r[k].y) + (r[k].z + r[k].w);
    }
    const double cnt = (double)(e1 - e0);
    const double mean_d = block_sum_d((double)s, red) / cnt;
    const float mean = (float)mean_d;
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < GN2_VEC; ++k) {
        const int64_t i = (int64_t)k * GN2_THREADS + threadIdx.x;
        if (i < nv) {
            const float a = r[k].x - mean, b = r[k].y - mean, c = r[k].z - mean, d = r[k].w - mean;
            s2 += (a * a + b * b) + (c * c + d * d);
        }
    }
    // sum of squared deviations from the ROUNDED mean; shift to the exact chunk mean: sum (x - m)^2 = sum (x - mf)^2 - cnt (m - mf)^2
    double m2 = block_sum_d((double)s2, red);
    const double dm = mean_d - (double)mean;
    m2 -= cnt * dm * dm;
    const int64_t ng = (int64_t)n * G + g;
    if (threadIdx.x == 0) {
        partial[ng * S + ci] = make_double2(mean_d, m2);
        __threadfence();
        const unsigned old = atomicAdd(&counters[ng], 1u);
        is_last = (old == (unsigned)(S - 1));
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        double na = 0.0, ma = 0.0, M2 = 0.0;
        for (int i = 0; i < S; ++i) {
            const double* pp = (const double*)partial + 2 * (ng * S + i);
            const double2 pb = make_double2(__ldcg(pp), __ldcg(pp + 1));     // written by other CTAs: read through L2
            const double nb = (double)(min((int64_t)(i + 1) * chunk, len) - (int64_t)i * chunk);
            const double nt = na + nb, delta = pb.x - ma;
            ma += delta * nb / nt;
            M2 += pb.y + delta * delta * na * nb / nt;
            na = nt;
        }
        const float var = (float)(M2 / na);
        stats[ng] = make_float2((float)ma, 1.0f / sqrtf(var + eps));
        counters[ng] = 0u;            // ready for the next launch (launches of one stream are ordered)
    }
}

// grid (ceil(OH*OW / (64 * NSUB)), C / 64, N), block 256: 64-channel x 64-pixel tiles through shared memory, NSUB of them per CTA.
// Load: 16 lanes x float4 cover 64 consecutive pixels of one channel (coalesced 256 B), per-channel norm/affine constants are
// fetched once per channel per thread; with NSUB = 4 a thread issues its 16 loads (1 KB of every channel row, 64 KB per CTA) before the
// first use -- the 64-pixel version kept only 16 KB per CTA in flight and ran at a third of the HBM rate on the VAE's 512x512 levels.
// Store: one warp writes one pixel's 64 channels = 128 contiguous bytes of the NHWC row.
template <int UP, int NSUB>
__global__ void __launch_bounds__(256, NSUB == 2 ? 4 : 1) k_to_nhwc_f16(const float* __restrict__ x, __half* __restrict__ out, const float2* __restrict__ stats,
                                                     const float* __restrict__ gw, const float* __restrict__ gb, int C, int H, int W, int OH,
                                                     int OW, int cpg, int G, int act, int vec_ok, const float* __restrict__ addv) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float tile[64][65];
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * 64;
    const int64_t OHW = (int64_t)OH * OW;
    const int64_t pbase = (int64_t)blockIdx.x * (64 * NSUB);
    const int t = threadIdx.x;
    const int px4 = (t & 15) * 4;          // first of 4 consecutive output pixels handled by this thread
    const int warp = t >> 5, lane = t & 31;
    float v[NSUB][4][4];
    // ---- all loads first
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
        const int64_t p0 = pbase + sub * 64;
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
            const int c = c0 + pass * 16 + (t >> 4);
            const float* xc = x + ((int64_t)n * C + c) * H * W;
            const int64_t p = p0 + px4;
            if (UP == 1 && vec_ok && p + 3 < OHW) {
                const float4 q = *(const float4*)(xc + p);     // OHW == H*W, 16-byte aligned: p % 4 == 0 and channel planes are multiples of 4 floats
                v[sub][pass][0] = q.x; v[sub][pass][1] = q.y; v[sub][pass][2] = q.z; v[sub][pass][3] = q.w;
            } else if (UP == 2 && vec_ok && (OW & 3) == 0 && p + 3 < OHW) {
                // nearest x2: four consecutive output pixels of one row (p % 4 == 0, OW % 4 == 0) are two source pixels, each twice -- one
                // 32-bit division and one 8-byte load instead of four 64-bit divisions and four loads (the pass was instruction bound)
                const unsigned pu = (unsigned)p, oy = pu / (unsigned)OW, ox = pu - oy * (unsigned)OW;
                const float2 q = *(const float2*)(xc + (int64_t)(oy >> 1) * W + (ox >> 1));
                v[sub][pass][0] = q.x; v[sub][pass][1] = q.x; v[sub][pass][2] = q.y; v[sub][pass][3] = q.y;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t pp = p + k;
                    if (pp < OHW) {
                        const int oy = (int)(pp / OW), ox = (int)(pp - (int64_t)oy * OW);
                        v[sub][pass][k] = xc[(int64_t)(oy / UP) * W + (ox / UP)];
                    } else {
                        v[sub][pass][k] = 0.f;
                    }
                }
            }
        }
    }
    // ---- per-channel constants (the same four channels for every sub-tile)
    float mean[4], rstd[4], w[4], b[4], pre[4];
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int c = c0 + pass * 16 + (t >> 4);
        mean[pass] = 0.f; rstd[pass] = 1.f; w[pass] = 1.f; b[pass] = 0.f;
        if (stats) { const float2 st = stats[(int64_t)n * G + c / cpg]; mean[pass] = st.x; rstd[pass] = st.y; }
        if (gw) { w[pass] = gw[c]; b[pass] = gb ? gb[c] : 0.f; }
        pre[pass] = addv ? addv[(int64_t)n * C + c] : 0.f;       // x + emb broadcast, added exactly where the unfused ADD rounds
    }
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
        const int64_t p0 = pbase + sub * 64;
        if (p0 >= OHW) break;
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
            const int cl = pass * 16 + (t >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float o = ((addv ? v[sub][pass][k] + pre[pass] : v[sub][pass][k]) - mean[pass]) * rstd[pass];
                o = o * w[pass] + b[pass];
                // SiLU with the fast exponential / division (MUFU.EX2, MUFU.RCP): the IEEE division and expf expanded to ~35 instructions per
                // element and made this pass INSTRUCTION bound (96 us for the 201 MB of a 512 x 512 x 128 VAE level, 2.1 TB/s); a few f32 ulp
                // of difference disappear in the rounding to f16 that follows (the CPU oracle's vectorised SiLU is a polynomial itself)
                if (act == 1) o = __fdividef(o, 1.0f + __expf(-o));
                tile[cl][px4 + k] = o;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int pl = warp * 8 + k;
            const int64_t p = p0 + pl;
            if (p < OHW) {
                const __half2 hv = __floats2half2_rn(tile[2 * lane][pl], tile[2 * lane + 1][pl]);
                *(__half2*)(out + ((int64_t)n * OHW + p) * C + c0 + 2 * lane) = hv;
            }
        }
        if (NSUB > 1) __syncthreads();
    }
}

__global__ void k_pack_conv_weight(const __half* __restrict__ w, __half* __restrict__ out, int KW, int KH, int IC, int64_t total) {
    pdl_wait();
    pdl_launch_dependents();
    // out[((oc * KH + kh) * KW + kw) * IC + ic] = w[((oc * IC + ic) * KH + kh) * KW + kw]
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int ic = (int)(r % IC); r /= IC;
        const int kw = (int)(r % KW); r /= KW;
        const int kh = (int)(r % KH);
        const int64_t oc = r / KH;
        out[i] = w[((oc * IC + ic) * KH + kh) * KW + kw];
    }
}

// one thread per Q8_0 block: 2-byte scale + 32 int8 (the block is only 2-byte aligned) -> 32 f16
__global__ void k_dequant_q8_0(const uint8_t* __restrict__ blocks, __half* __restrict__ out, int64_t n_blocks) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t* blk = blocks + b * 34;
        const float d = __half2float(*(const __half*)blk);
        const uint16_t* q16 = (const uint16_t*)(blk + 2);
        __half2* o = (__half2*)(out + b * 32);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint16_t two = q16[i];
            const float a = d * (float)(int8_t)(two & 0xff), c = d * (float)(int8_t)(two >> 8);
            o[i] = __floats2half2_rn(a, c);
        }
    }
}

}  // namespace

int b200_launch_dequant_q8_0(cudaStream_t s, const void* blocks, void* out_f16, int64_t n_blocks) {
    if (n_blocks <= 0) return 0;
    int64_t nb = (n_blocks + 255) / 256;
    if (nb > 148 * 64) nb = 148 * 64;
    b200_launch(k_dequant_q8_0, dim3((unsigned)nb), dim3(256), 0, s, (const uint8_t*)blocks, (__half*)out_f16, n_blocks);
    return 1;
}

int b200_launch_gn_stats(cudaStream_t s, const float* x, float* stats, int64_t N, int64_t C, int64_t inner, int n_groups, float eps, void* partial,
                         unsigned* counters, const float* addv) {
    const int cpg = (int)((C + n_groups - 1) / n_groups);
    const int64_t len = (int64_t)cpg * inner;
    const int64_t S = (len + GN2_CHUNK - 1) / GN2_CHUNK;
    // chunked path: every group complete (C % cpg == 0), float4-aligned chunks, counters available
    if (partial && counters && S >= 2 && S <= 4096 && C % cpg == 0 && (inner & 3) == 0 && (((uintptr_t)x) & 15) == 0 && N * n_groups <= B200_GN_COUNTERS && N <= 65535) {
        int64_t chunk = (len + S - 1) / S;
        chunk = (chunk + 3) & ~(int64_t)3;
        dim3 grid((unsigned)S, (unsigned)n_groups, (unsigned)N);
        b200_launch(k_gn_stats_chunked, dim3(grid), dim3(GN2_THREADS), 0, s, x, (float2*)stats, (double2*)partial, counters, inner, (int)C, cpg, n_groups, (int)S,
                    chunk, eps, addv);
        return 1;
    }
    dim3 grid((unsigned)n_groups, (unsigned)N);
    b200_launch(k_gn_stats, dim3(grid), dim3(1024), 0, s, x, (float2*)stats, inner, (int)C, cpg, n_groups, eps, addv);
    return 1;
}

size_t b200_gn_stats_partial_bytes(int64_t N, int64_t C, int64_t inner, int n_groups) {
    const int cpg = (int)((C + n_groups - 1) / n_groups);
    const int64_t S = ((int64_t)cpg * inner + GN2_CHUNK - 1) / GN2_CHUNK;
    return (size_t)(N * n_groups * S * 16);
}

int b200_launch_to_nhwc_f16(cudaStream_t s, const float* x, void* out, int64_t N, int64_t C, int64_t H, int64_t W, int up, const float* stats,
                            int n_groups, const float* gw, const float* gb, int act, const float* addv) {
    if (C % 64 != 0) return -1;
    const int64_t OH = H * up, OW = W * up;
    const int cpg = n_groups > 0 ? (int)((C + n_groups - 1) / n_groups) : 1;
    // large images: 128 pixels per CTA, four CTAs per SM (32 KB of loads in flight each, four phases that overlap).  Measured on B200 against
    // 256 pixels / two CTAs per SM (GGML_B200_NHWC_NSUB=4): VAE decode 5.61 -> 5.29 ms, 1024 x 1024 decode 26.2 -> 25.2 ms
    static int nsub_big = -1;
    if (nsub_big < 0) { const char* e = getenv("GGML_B200_NHWC_NSUB"); nsub_big = (e && atoi(e) == 4) ? 4 : 2; }
    const int nsub = (up == 1 && OH * OW >= 16384) ? nsub_big : 1;
    dim3 grid((unsigned)((OH * OW + 64 * nsub - 1) / (64 * nsub)), (unsigned)(C / 64), (unsigned)N);
    if (grid.y > 65535 || N > 65535 || (up != 1 && up != 2)) return -1;
    const int vec_ok = (((uintptr_t)x & 15) == 0 && ((H * W) & 3) == 0) ? 1 : 0;
#define NHWC_ARGS x, (__half*)out, (const float2*)stats, gw, gb, (int)C, (int)H, (int)W, (int)OH, (int)OW, cpg, n_groups, act, vec_ok, addv
    if (up == 1 && nsub == 4) b200_launch(k_to_nhwc_f16<1, 4>, dim3(grid), dim3(256), 0, s, NHWC_ARGS);
    else if (up == 1 && nsub == 2) b200_launch(k_to_nhwc_f16<1, 2>, dim3(grid), dim3(256), 0, s, NHWC_ARGS);
    else if (up == 1) b200_launch(k_to_nhwc_f16<1, 1>, dim3(grid), dim3(256), 0, s, NHWC_ARGS);
    else b200_launch(k_to_nhwc_f16<2, 1>, dim3(grid), dim3(256), 0, s, NHWC_ARGS);
#undef NHWC_ARGS
    return 1;
}

int b200_launch_pack_conv_weight(cudaStream_t s, const void* w, void* out, int KW, int KH, int64_t IC, int64_t OC) {
    const int64_t total = (int64_t)KW * KH * IC * OC;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    b200_launch(k_pack_conv_weight, dim3((unsigned)blocks), dim3(256), 0, s, (const __half*)w, (__half*)out, KW, KH, (int)IC, total);
    return 1;
}
