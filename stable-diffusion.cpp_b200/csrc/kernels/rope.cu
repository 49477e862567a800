// rope.cu -- the reference's rotary embedding as ONE pass (Rope::apply_rope, src/model/common/rope.hpp:966-1010, interleaved
// variant; Flux double/single-stream blocks flux.hpp:279-295,493-699 and the Wan attention blocks).
//
// The reference graph spends 8 kernels and ~1 GB of HBM traffic per q or k tensor of a FLUX.1 block on it (CONT of the permute, CONT
// of the even/odd split, two REPEATs, CONT of pe, two MULs, ADD).  Here each warp owns one (token, head) row of d_head floats:
//   [optional RMSNorm over the row * per-channel scale  (QKNorm, flux.hpp:213-261)]
//   out[2i + j] = x[2i] * pe[a=0, b=j, i, l] + x[2i+1] * pe[a=1, b=j, i, l]        pe: ggml [2(a), 2(b), d/2, L] = [[cos, -sin], [sin, cos]]
// written as f32 or f16 rows in [d, L, head] order (what ggml_ext_attention_ext consumes with skip_reshape).  The two products are
// rounded separately and then added (no FMA contraction), exactly like the unfused MUL, MUL, ADD nodes.
// HBM roofline: algorithmic bytes = rows * d * (4 + 4|2) + |pe|.
#include "../b200_ops.h"
#include "b200_launch.cuh"

#include <cuda_fp16.h>

namespace {

template <typename TD>
__global__ void __launch_bounds__(256) k_rope_rows(const char* __restrict__ x, int64_t nb_h, int64_t nb_l, int64_t nb_n, const float* __restrict__ pe,
                                                   TD* __restrict__ out, int d, int H, int L, int64_t rows, const float* __restrict__ rms_w, float eps,
                                                   int has_rms) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        // output row order: l fastest, then head, then batch  (out is [d, L, H, N])
        const int l = (int)(row % L);
        const int64_t t = row / L;
        const int h = (int)(t % H);
        const int64_t n = t / H;
        const float* xr = (const float*)(x + h * nb_h + l * nb_l + n * nb_n);
        float rs = 1.0f;
        if (has_rms) {
            float ss = 0.f;
            for (int c = lane * 4; c < d; c += 128) {
                const float4 v = *(const float4*)(xr + c);
                ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            rs = 1.0f / sqrtf(ss / (float)d + eps);
        }
        TD* orow = out + row * d;
        const float4* per = (const float4*)(pe + (int64_t)l * (d / 2) * 4);
        for (int c = lane * 4; c < d; c += 128) {
            float4 v = *(const float4*)(xr + c);
            if (has_rms) {
                const float4 w = *(const float4*)(rms_w + c);
                v.x = __fmul_rn(__fmul_rn(v.x, rs), w.x); v.y = __fmul_rn(__fmul_rn(v.y, rs), w.y);
                v.z = __fmul_rn(__fmul_rn(v.z, rs), w.z); v.w = __fmul_rn(__fmul_rn(v.w, rs), w.w);
            }
            const float4 p0 = per[c / 2], p1 = per[c / 2 + 1];     // (a0b0, a1b0, a0b1, a1b1) of pair i = c/2 and i + 1
            const float o0 = __fadd_rn(__fmul_rn(v.x, p0.x), __fmul_rn(v.y, p0.y));
            const float o1 = __fadd_rn(__fmul_rn(v.x, p0.z), __fmul_rn(v.y, p0.w));
            const float o2 = __fadd_rn(__fmul_rn(v.z, p1.x), __fmul_rn(v.w, p1.y));
            const float o3 = __fadd_rn(__fmul_rn(v.z, p1.z), __fmul_rn(v.w, p1.w));
            if (sizeof(TD) == 4) {
                *(float4*)((float*)orow + c) = make_float4(o0, o1, o2, o3);
            } else {
                const __half2 a = __floats2half2_rn(o0, o1), b = __floats2half2_rn(o2, o3);
                uint2 u;
                u.x = *(const uint32_t*)&a;
                u.y = *(const uint32_t*)&b;
                *(uint2*)((__half*)orow + c) = u;
            }
        }
    }
}

}  // namespace

// x: f32 view [d, H, L, N] (rows unit-stride, 16-byte aligned; strides in bytes), pe: contiguous f32 [2, 2, d/2, L], out: contiguous
// [d, L, H, N] of out_type (GGML_TYPE_F32 | GGML_TYPE_F16).  rms_w != null: RMSNorm(eps) over each row times rms_w[d] first.
int b200_launch_rope(cudaStream_t s, const b200_td& x, const float* pe, void* out, int out_type, const float* rms_w, float eps) {
    const int64_t d = x.ne[0], H = x.ne[1], L = x.ne[2], N = x.ne[3];
    if (x.type != GGML_TYPE_F32 || x.nb[0] != 4 || d % 4 || d < 4 || ((uintptr_t)x.data & 15) || (x.nb[1] & 15) || (x.nb[2] & 15) || (x.nb[3] & 15)) return -1;
    if (((uintptr_t)pe & 15) || ((uintptr_t)out & 15) || (rms_w && ((uintptr_t)rms_w & 15))) return -1;
    if (out_type != GGML_TYPE_F32 && out_type != GGML_TYPE_F16) return -1;
    if (H > 0x7fffffff || L > 0x7fffffff || d > 0x7fffffff) return -1;
    const int64_t rows = H * L * N;
    if (rows == 0) return 0;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (out_type == GGML_TYPE_F32)
        b200_launch(k_rope_rows<float>, dim3((unsigned)blocks), dim3(256), 0, s, (const char*)x.data, x.nb[1], x.nb[2], x.nb[3], pe, (float*)out, (int)d, (int)H, (int)L,
                    rows, rms_w, eps, rms_w ? 1 : 0);
    else
        b200_launch(k_rope_rows<__half>, dim3((unsigned)blocks), dim3(256), 0, s, (const char*)x.data, x.nb[1], x.nb[2], x.nb[3], pe, (__half*)out, (int)d, (int)H,
                    (int)L, rows, rms_w, eps, rms_w ? 1 : 0);
    return 1;
}
