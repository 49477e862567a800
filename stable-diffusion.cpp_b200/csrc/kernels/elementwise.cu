// elementwise.cu -- HBM-bound data-movement and pointwise kernels of the B200 ggml backend.
//
// Semantics follow the reference's CPU backend (the oracle):
//   binary bcast   ggml/src/ggml-cpu/binary-ops.cpp          unary   ggml/src/ggml-cpu/unary-ops.cpp, vec.h:963-1060
//   cpy/cont/dup   ggml/src/ggml-cpu/ops.cpp (dup_*)          concat  ops.cpp (concat_f32)
//   upscale        ops.cpp:7832                               pad     ops.cpp (pad_f32)
//   timestep_emb   ops.cpp:8278-8309                          repeat  ops.cpp (repeat_f32)
// Roofline: all of these are pure bandwidth (<= 1 flop/byte): grids are sized in multiples of the SM
// count with 16-byte accesses on the contiguous fast paths; the strided generic paths exist for
// coverage (test-backend-ops) and are replaced by fused producers/consumers in whole-model graphs.
#include "../b200_ops.h"
#include "b200_launch.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace {

constexpr int kThreads = 256;

__host__ __device__ inline int64_t td_nelements(const b200_td& t) { return t.ne[0] * t.ne[1] * t.ne[2] * t.ne[3]; }

inline bool td_contiguous(const b200_td& t, int64_t esize) {
    int64_t s = esize;
    for (int i = 0; i < 4; ++i) {
        if (t.ne[i] != 1 && t.nb[i] != s) return false;
        s *= t.ne[i];
    }
    return true;
}
inline int64_t type_size(int type) {
    switch (type) {
        case GGML_TYPE_F32: case GGML_TYPE_I32: return 4;
        case GGML_TYPE_F16: case GGML_TYPE_BF16: case GGML_TYPE_I16: return 2;
        case GGML_TYPE_I8: return 1;
        case GGML_TYPE_I64: case GGML_TYPE_F64: return 8;
        default: return 0;
    }
}
inline unsigned grid_for(int64_t n, int per_thread = 1) {
    int64_t b = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    if (b < 1) b = 1;
    if (b > 0x7fffffff) b = 0x7fffffff;
    return (unsigned)b;
}

template <typename T> __device__ __forceinline__ float ldf(const void* p);
template <> __device__ __forceinline__ float ldf<float>(const void* p) { return *(const float*)p; }
template <> __device__ __forceinline__ float ldf<__half>(const void* p) { return __half2float(*(const __half*)p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const void* p) { return __bfloat162float(*(const __nv_bfloat16*)p); }
template <typename T> __device__ __forceinline__ void stf(void* p, float v);
template <> __device__ __forceinline__ void stf<float>(void* p, float v) { *(float*)p = v; }
template <> __device__ __forceinline__ void stf<__half>(void* p, float v) { *(__half*)p = __float2half_rn(v); }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(void* p, float v) { *(__nv_bfloat16*)p = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------
// binary broadcast
// ---------------------------------------------------------------------------------------------
template <int OP> __device__ __forceinline__ float binop(float a, float b) {
    if (OP == B200_ADD) return a + b;
    if (OP == B200_SUB) return a - b;
    if (OP == B200_MUL) return a * b;
    return a / b;
}

// ---------------------------------------------------------------------------------------------
// row-vectorised fast paths: tensors whose rows (dim 0) are unit-stride and 16-byte aligned but whose outer dims are
// strided (views, permutes).  One thread per 4 elements; the (i1,i2,i3) decomposition is 32-bit and per 16 bytes, not
// 64-bit per element like the generic kernels.
// ---------------------------------------------------------------------------------------------
struct rows4 {
    uint32_t cpr;            // 4-element chunks per row
    uint32_t ne1, ne2;       // row -> (i1, i2, i3)
    uint32_t total;          // chunks
};
struct strides3 { int64_t nb1, nb2, nb3; };

__device__ __forceinline__ void rows4_decode(const rows4& r, uint32_t idx, uint32_t& c, uint32_t& i1, uint32_t& i2, uint32_t& i3) {
    uint32_t row = idx / r.cpr;
    c = idx - row * r.cpr;
    uint32_t t = row / r.ne1;
    i1 = row - t * r.ne1;
    i3 = t / r.ne2;
    i2 = t - i3 * r.ne2;
}

template <int OP>
__global__ void k_binary_rows4(const char* __restrict__ a, strides3 sa, const char* __restrict__ b, strides3 sb, char* __restrict__ d, strides3 sd,
                               rows4 r) {
    pdl_wait();
    pdl_launch_dependents();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < r.total; idx += gridDim.x * blockDim.x) {
        uint32_t c, i1, i2, i3;
        rows4_decode(r, idx, c, i1, i2, i3);
        const float4 x = *(const float4*)(a + i1 * sa.nb1 + i2 * sa.nb2 + i3 * sa.nb3 + (int64_t)c * 16);
        const float4 y = *(const float4*)(b + i1 * sb.nb1 + i2 * sb.nb2 + i3 * sb.nb3 + (int64_t)c * 16);
        *(float4*)(d + i1 * sd.nb1 + i2 * sd.nb2 + i3 * sd.nb3 + (int64_t)c * 16) =
            make_float4(binop<OP>(x.x, y.x), binop<OP>(x.y, y.y), binop<OP>(x.z, y.z), binop<OP>(x.w, y.w));
    }
}

__global__ void k_copy_rows4(const char* __restrict__ a, strides3 sa, char* __restrict__ d, strides3 sd, rows4 r) {
    pdl_wait();
    pdl_launch_dependents();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < r.total; idx += gridDim.x * blockDim.x) {
        uint32_t c, i1, i2, i3;
        rows4_decode(r, idx, c, i1, i2, i3);
        *(uint4*)(d + i1 * sd.nb1 + i2 * sd.nb2 + i3 * sd.nb3 + (int64_t)c * 16) =
            *(const uint4*)(a + i1 * sa.nb1 + i2 * sa.nb2 + i3 * sa.nb3 + (int64_t)c * 16);
    }
}

// concat along dim 0 (the single-stream Flux block glues attention and MLP halves of every token row, flux.hpp:690): destination
// chunk -> which source row segment, then a 16-byte copy
__global__ void k_concat_dim0_rows4(const char* __restrict__ a, strides3 sa, const char* __restrict__ b, strides3 sb, char* __restrict__ d, strides3 sd,
                                    rows4 r, uint32_t a_chunks) {
    pdl_wait();
    pdl_launch_dependents();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < r.total; idx += gridDim.x * blockDim.x) {
        uint32_t c, i1, i2, i3;
        rows4_decode(r, idx, c, i1, i2, i3);
        const uint4 v = c < a_chunks ? *(const uint4*)(a + i1 * sa.nb1 + i2 * sa.nb2 + i3 * sa.nb3 + (int64_t)c * 16)
                                     : *(const uint4*)(b + i1 * sb.nb1 + i2 * sb.nb2 + i3 * sb.nb3 + (int64_t)(c - a_chunks) * 16);
        *(uint4*)(d + i1 * sd.nb1 + i2 * sd.nb2 + i3 * sd.nb3 + (int64_t)c * 16) = v;
    }
}

// GEGLU tail of the reference's FeedForward (src/model/common/block.hpp:194-207): dst = x * gelu_tanh(gate), x and gate the two
// halves (strided row views) of one projection output.  Replaces CONT(gate) + GELU + MUL (+ the f16 operand pack of the Linear
// that follows, through the optional contiguous f16 shadow d16).  Arithmetic is the unfused kernels' own: f32 gelu, one multiply.
__device__ __forceinline__ float gelu_tanh_f32(float x) {
    return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
}
__global__ void k_geglu_rows4(const char* __restrict__ a, strides3 sa, const char* __restrict__ b, strides3 sb, char* __restrict__ d, strides3 sd,
                              __half* __restrict__ d16, rows4 r) {
    pdl_wait();
    pdl_launch_dependents();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < r.total; idx += gridDim.x * blockDim.x) {
        uint32_t c, i1, i2, i3;
        rows4_decode(r, idx, c, i1, i2, i3);
        const float4 x = *(const float4*)(a + i1 * sa.nb1 + i2 * sa.nb2 + i3 * sa.nb3 + (int64_t)c * 16);
        const float4 g = *(const float4*)(b + i1 * sb.nb1 + i2 * sb.nb2 + i3 * sb.nb3 + (int64_t)c * 16);
        const float4 y = make_float4(x.x * gelu_tanh_f32(g.x), x.y * gelu_tanh_f32(g.y), x.z * gelu_tanh_f32(g.z), x.w * gelu_tanh_f32(g.w));
        *(float4*)(d + i1 * sd.nb1 + i2 * sd.nb2 + i3 * sd.nb3 + (int64_t)c * 16) = y;
        if (d16) {   // dst is contiguous: chunk idx is flat element 4 * idx
            const __half2 lo = __floats2half2_rn(y.x, y.y), hi = __floats2half2_rn(y.z, y.w);
            uint2 u;
            u.x = *(const uint32_t*)&lo;
            u.y = *(const uint32_t*)&hi;
            *(uint2*)(d16 + (size_t)idx * 4) = u;
        }
    }
}

// concat along dim >= 1 of row-contiguous 4-byte tensors: destination row -> which source, then a 16-byte copy
__global__ void k_concat_rows4(const char* __restrict__ a, strides3 sa, const char* __restrict__ b, strides3 sb, char* __restrict__ d, strides3 sd,
                               rows4 r, int dim, uint32_t a_ne_dim) {
    pdl_wait();
    pdl_launch_dependents();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < r.total; idx += gridDim.x * blockDim.x) {
        uint32_t c, i[4];
        rows4_decode(r, idx, c, i[1], i[2], i[3]);
        char* pd = d + i[1] * sd.nb1 + i[2] * sd.nb2 + i[3] * sd.nb3 + (int64_t)c * 16;
        const char* src = a;
        strides3 ss = sa;
        if (i[dim] >= a_ne_dim) { i[dim] -= a_ne_dim; src = b; ss = sb; }
        *(uint4*)pd = *(const uint4*)(src + i[1] * ss.nb1 + i[2] * ss.nb2 + i[3] * ss.nb3 + (int64_t)c * 16);
    }
}

// f32 rows -> dense f16/bf16 [rows][kpad], 8 outputs (16 bytes) per thread; rows unit-stride and 16-byte aligned
template <typename TD>
__global__ void k_pack_rows8(const char* __restrict__ a, strides3 sa, TD* __restrict__ d, rows4 r /* cpr = kpad/8 */, uint32_t K) {
    pdl_wait();
    pdl_launch_dependents();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < r.total; idx += gridDim.x * blockDim.x) {
        uint32_t c, i1, i2, i3;
        rows4_decode(r, idx, c, i1, i2, i3);
        const float* src = (const float*)(a + i1 * sa.nb1 + i2 * sa.nb2 + i3 * sa.nb3) + c * 8;
        float v[8];
        if (c * 8 + 8 <= K) {
            const float4 x = *(const float4*)src, y = *(const float4*)(src + 4);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (c * 8 + k < K) ? src[k] : 0.f;
        }
        TD o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) stf<TD>(&o[k], v[k]);
        *(uint4*)(d + (int64_t)idx * 8) = *(const uint4*)o;
    }
}

inline bool rows4_ok(const b200_td& t) {
    return t.nb[0] == 4 && t.ne[0] % 4 == 0 && (uintptr_t)t.data % 16 == 0 && t.nb[1] % 16 == 0 && t.nb[2] % 16 == 0 && t.nb[3] % 16 == 0;
}
inline bool make_rows4(const b200_td& shape, rows4* r) {
    int64_t total = (shape.ne[0] / 4) * shape.ne[1] * shape.ne[2] * shape.ne[3];
    if (total <= 0 || total >= (1ll << 31) || shape.ne[0] / 4 >= (1ll << 31) || shape.ne[1] >= (1ll << 31) || shape.ne[2] >= (1ll << 31)) return false;
    r->cpr = (uint32_t)(shape.ne[0] / 4); r->ne1 = (uint32_t)shape.ne[1]; r->ne2 = (uint32_t)shape.ne[2]; r->total = (uint32_t)total;
    return true;
}
inline strides3 st3(const b200_td& t) { return strides3{t.nb[1], t.nb[2], t.nb[3]}; }
inline bool same_shape(const b200_td& a, const b200_td& b) { return a.ne[0] == b.ne[0] && a.ne[1] == b.ne[1] && a.ne[2] == b.ne[2] && a.ne[3] == b.ne[3]; }


// generic: any strides, src1 broadcast by modulo (ggml_can_repeat(src1, src0))
template <int OP, typename TA, typename TB, typename TD>
__global__ void k_binary_generic(b200_td a, b200_td b, b200_td d, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = i % d.ne[0], r = i / d.ne[0];
        int64_t i1 = r % d.ne[1];
        r /= d.ne[1];
        int64_t i2 = r % d.ne[2], i3 = r / d.ne[2];
        const char* pa = (const char*)a.data + i0 * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3];
        const char* pb = (const char*)b.data + (i0 % b.ne[0]) * b.nb[0] + (i1 % b.ne[1]) * b.nb[1] + (i2 % b.ne[2]) * b.nb[2] + (i3 % b.ne[3]) * b.nb[3];
        char* pd = (char*)d.data + i0 * d.nb[0] + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3];
        stf<TD>(pd, binop<OP>(ldf<TA>(pa), ldf<TB>(pb)));
    }
}

// fast path: a, d contiguous f32 with identical shape; b contiguous f32 and either
//   mode 0: same shape, mode 1: b = [ne0,1,1,1] (row vector), mode 2: b = [1,1,C,1|N] over inner = ne0*ne1 (channel vector, optionally per image)
template <int OP>
__global__ void k_binary_f32_vec4(const float4* __restrict__ a, const float* __restrict__ b, float4* __restrict__ d, int64_t n4,
                                  int mode, int64_t ne0, int64_t inner, int64_t C) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = a[i], y;
        if (mode == 0) {
            y = ((const float4*)b)[i];
        } else if (mode == 1) {
            y = *(const float4*)(b + (i * 4) % ne0);
        } else {
            float v = b[((i * 4) / inner) % C];
            y = make_float4(v, v, v, v);
        }
        d[i] = make_float4(binop<OP>(x.x, y.x), binop<OP>(x.y, y.y), binop<OP>(x.z, y.z), binop<OP>(x.w, y.w));
    }
}

template <int OP>
int launch_binary_op(cudaStream_t s, const b200_td& a, const b200_td& b, const b200_td& d) {
    int64_t n = td_nelements(d);
    if (n == 0) return 0;
    bool f32 = a.type == GGML_TYPE_F32 && b.type == GGML_TYPE_F32 && d.type == GGML_TYPE_F32;
    if (f32 && td_contiguous(a, 4) && td_contiguous(d, 4) && td_contiguous(b, 4) && n % 4 == 0 &&
        ((uintptr_t)a.data % 16 == 0) && ((uintptr_t)d.data % 16 == 0) && ((uintptr_t)b.data % 16 == 0)) {
        int mode = -1;
        int64_t nb_el = td_nelements(b);
        if (nb_el == n) mode = 0;
        else if (nb_el == b.ne[0] && b.ne[0] == d.ne[0] && d.ne[0] % 4 == 0) mode = 1;
        // channel vector [1,1,C,1] over every image, or one vector per image [1,1,C,N] (the `h + emb` of a batched-CFG ResBlock): a and d are
        // contiguous [W,H,C,N], so element i belongs to flat channel i / (W*H) = c + C*n, which is b's own flat index
        else if (b.ne[0] == 1 && b.ne[1] == 1 && b.ne[2] == d.ne[2] && (b.ne[3] == 1 || b.ne[3] == d.ne[3]) && (d.ne[0] * d.ne[1]) % 4 == 0) mode = 2;
        if (mode >= 0) {
            b200_launch(k_binary_f32_vec4<OP>, dim3(grid_for(n / 4)), dim3(kThreads), 0, s, (const float4*)a.data, (const float*)b.data, (float4*)d.data, n / 4, mode,
                                                                        d.ne[0], d.ne[0] * d.ne[1], b.ne[2] * b.ne[3]);
            return 1;
        }
    }
    if (f32 && same_shape(a, d) && same_shape(b, d) && rows4_ok(a) && rows4_ok(b) && rows4_ok(d)) {
        rows4 r;
        if (make_rows4(d, &r)) {
            b200_launch(k_binary_rows4<OP>, dim3(grid_for(r.total)), dim3(kThreads), 0, s, (const char*)a.data, st3(a), (const char*)b.data, st3(b), (char*)d.data, st3(d), r);
            return 1;
        }
    }
    unsigned g = grid_for(n);
#define BIN_CASE(TA, TB, TD) b200_launch(k_binary_generic<OP, TA, TB, TD>, dim3(g), dim3(kThreads), 0, s, a, b, d, n)
    if (f32) BIN_CASE(float, float, float);
    else if (a.type == GGML_TYPE_F16 && b.type == GGML_TYPE_F16 && d.type == GGML_TYPE_F16) BIN_CASE(__half, __half, __half);
    else if (a.type == GGML_TYPE_F32 && b.type == GGML_TYPE_F16 && d.type == GGML_TYPE_F32) BIN_CASE(float, __half, float);
    else if (a.type == GGML_TYPE_F16 && b.type == GGML_TYPE_F32 && d.type == GGML_TYPE_F32) BIN_CASE(__half, float, float);
    else if (a.type == GGML_TYPE_F16 && b.type == GGML_TYPE_F32 && d.type == GGML_TYPE_F16) BIN_CASE(__half, float, __half);
    else if (a.type == GGML_TYPE_BF16 && b.type == GGML_TYPE_BF16 && d.type == GGML_TYPE_BF16) BIN_CASE(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
    else return -1;
#undef BIN_CASE
    return 1;
}

// ---------------------------------------------------------------------------------------------
// unary / scalar ops
// ---------------------------------------------------------------------------------------------
struct UnaryParams { int op; float p0, p1; };

// op codes >= 100 are b200_scalar_op + 100
__device__ __forceinline__ float apply_unary(int op, float x, float p0, float p1) {
    switch (op) {
        case GGML_UNARY_OP_ABS: return fabsf(x);
        case GGML_UNARY_OP_SGN: return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
        case GGML_UNARY_OP_NEG: return -x;
        case GGML_UNARY_OP_STEP: return x > 0.f ? 1.f : 0.f;
        case GGML_UNARY_OP_TANH: return tanhf(x);
        case GGML_UNARY_OP_ELU: return x > 0.f ? x : expm1f(x);
        case GGML_UNARY_OP_RELU: return fmaxf(x, 0.f);
        case GGML_UNARY_OP_SIGMOID: return 1.f / (1.f + expf(-x));
        case GGML_UNARY_OP_GELU: return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
        case GGML_UNARY_OP_GELU_QUICK: return x * (1.0f / (1.0f + expf(-1.702f * x)));
        case GGML_UNARY_OP_SILU: return x / (1.0f + expf(-x));
        case GGML_UNARY_OP_HARDSWISH: return x * fminf(1.f, fmaxf(0.f, (x + 3.f) / 6.f));
        case GGML_UNARY_OP_HARDSIGMOID: return fminf(1.f, fmaxf(0.f, (x + 3.f) / 6.f));
        case GGML_UNARY_OP_EXP: return expf(x);
        case GGML_UNARY_OP_EXPM1: return expm1f(x);
        case GGML_UNARY_OP_SOFTPLUS: return x > 20.f ? x : log1pf(expf(x));
        case GGML_UNARY_OP_GELU_ERF: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
        case GGML_UNARY_OP_FLOOR: return floorf(x);
        case GGML_UNARY_OP_CEIL: return ceilf(x);
        case GGML_UNARY_OP_ROUND: return roundf(x);
        case GGML_UNARY_OP_TRUNC: return truncf(x);
        case 100 + B200_SCALE: return x * p0 + p1;
        case 100 + B200_CLAMP: return fminf(fmaxf(x, p0), p1);
        case 100 + B200_SQR: return x * x;
        case 100 + B200_SQRT: return sqrtf(x);
        case 100 + B200_LEAKY_RELU: return x > 0.f ? x : x * p0;
        case 100 + B200_SIN: return sinf(x);
        case 100 + B200_COS: return cosf(x);
        case 100 + B200_LOG: return logf(x);
    }
    return x;
}

template <typename TS, typename TD>
__global__ void k_unary_generic(b200_td a, b200_td d, int64_t n, UnaryParams p) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = i % d.ne[0], r = i / d.ne[0];
        int64_t i1 = r % d.ne[1];
        r /= d.ne[1];
        int64_t i2 = r % d.ne[2], i3 = r / d.ne[2];
        const char* pa = (const char*)a.data + i0 * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3];
        char* pd = (char*)d.data + i0 * d.nb[0] + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3];
        stf<TD>(pd, apply_unary(p.op, ldf<TS>(pa), p.p0, p.p1));
    }
}

__global__ void k_unary_f32_vec4(const float4* __restrict__ a, float4* __restrict__ d, int64_t n4, UnaryParams p) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = a[i];
        d[i] = make_float4(apply_unary(p.op, x.x, p.p0, p.p1), apply_unary(p.op, x.y, p.p0, p.p1), apply_unary(p.op, x.z, p.p0, p.p1),
                           apply_unary(p.op, x.w, p.p0, p.p1));
    }
}

int launch_unary_impl(cudaStream_t s, const b200_td& a, const b200_td& d, UnaryParams p) {
    int64_t n = td_nelements(d);
    if (n == 0) return 0;
    if (a.type == GGML_TYPE_F32 && d.type == GGML_TYPE_F32 && td_contiguous(a, 4) && td_contiguous(d, 4) && n % 4 == 0 &&
        (uintptr_t)a.data % 16 == 0 && (uintptr_t)d.data % 16 == 0) {
        b200_launch(k_unary_f32_vec4, dim3(grid_for(n / 4)), dim3(kThreads), 0, s, (const float4*)a.data, (float4*)d.data, n / 4, p);
        return 1;
    }
    unsigned g = grid_for(n);
    if (a.type == GGML_TYPE_F32 && d.type == GGML_TYPE_F32) b200_launch(k_unary_generic<float, float>, dim3(g), dim3(kThreads), 0, s, a, d, n, p);
    else if (a.type == GGML_TYPE_F16 && d.type == GGML_TYPE_F16) b200_launch(k_unary_generic<__half, __half>, dim3(g), dim3(kThreads), 0, s, a, d, n, p);
    else return -1;
    return 1;
}

// ---------------------------------------------------------------------------------------------
// copy / cont / dup (type converting, arbitrary strides, shapes may differ but element counts match)
// ---------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void k_copy_generic(b200_td a, b200_td d, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        int64_t a0 = r % a.ne[0]; r /= a.ne[0];
        int64_t a1 = r % a.ne[1]; r /= a.ne[1];
        int64_t a2 = r % a.ne[2]; int64_t a3 = r / a.ne[2];
        r = i;
        int64_t d0 = r % d.ne[0]; r /= d.ne[0];
        int64_t d1 = r % d.ne[1]; r /= d.ne[1];
        int64_t d2 = r % d.ne[2]; int64_t d3 = r / d.ne[2];
        const char* pa = (const char*)a.data + a0 * a.nb[0] + a1 * a.nb[1] + a2 * a.nb[2] + a3 * a.nb[3];
        char* pd = (char*)d.data + d0 * d.nb[0] + d1 * d.nb[1] + d2 * d.nb[2] + d3 * d.nb[3];
        stf<TD>(pd, ldf<TS>(pa));
    }
}

template <typename T>
__global__ void k_copy_raw(b200_td a, b200_td d, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();   // same-size raw element copy (ints, same float types)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        int64_t a0 = r % a.ne[0]; r /= a.ne[0];
        int64_t a1 = r % a.ne[1]; r /= a.ne[1];
        int64_t a2 = r % a.ne[2]; int64_t a3 = r / a.ne[2];
        r = i;
        int64_t d0 = r % d.ne[0]; r /= d.ne[0];
        int64_t d1 = r % d.ne[1]; r /= d.ne[1];
        int64_t d2 = r % d.ne[2]; int64_t d3 = r / d.ne[2];
        *(T*)((char*)d.data + d0 * d.nb[0] + d1 * d.nb[1] + d2 * d.nb[2] + d3 * d.nb[3]) =
            *(const T*)((const char*)a.data + a0 * a.nb[0] + a1 * a.nb[1] + a2 * a.nb[2] + a3 * a.nb[3]);
    }
}

// Tiled transpose for f32: dst contiguous; src has its unit-stride axis at dim `ax` != 0 of the SAME logical shape.
// Treats the tensor as batches of a 2-D problem: rows = dim ax (contiguous in src), cols = dim 0 (contiguous in dst).
// Requirements: src.nb[ax] == 4; shapes identical; the other two dims are looped over by blockIdx.z.
__global__ void k_transpose_f32(b200_td a, b200_td d, int ax, int o1, int o2) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float tile[32][33];
    int64_t z = blockIdx.z;
    int64_t j1 = z % d.ne[o1], j2 = z / d.ne[o1];
    const char* abase = (const char*)a.data + j1 * a.nb[o1] + j2 * a.nb[o2];
    char* dbase = (char*)d.data + j1 * d.nb[o1] + j2 * d.nb[o2];
    int64_t c0 = (int64_t)blockIdx.x * 32;   // along dim 0 (dst-contiguous)
    int64_t r0 = (int64_t)blockIdx.y * 32;   // along dim ax (src-contiguous)
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int64_t c = c0 + k, r = r0 + threadIdx.x;
        if (c < d.ne[0] && r < d.ne[ax]) tile[k][threadIdx.x] = *(const float*)(abase + c * a.nb[0] + r * a.nb[ax]);
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int64_t r = r0 + k, c = c0 + threadIdx.x;
        if (c < d.ne[0] && r < d.ne[ax]) *(float*)(dbase + c * d.nb[0] + r * d.nb[ax]) = tile[threadIdx.x][k];
    }
}

__global__ void k_copy_contig16(const uint4* __restrict__ a, uint4* __restrict__ d, int64_t n16) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) d[i] = a[i];
}

template <typename TD>
__global__ void k_cvt_f32_contig(const float4* __restrict__ a, TD* __restrict__ d, int64_t n4) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = a[i];
        stf<TD>(d + 4 * i + 0, x.x); stf<TD>(d + 4 * i + 1, x.y); stf<TD>(d + 4 * i + 2, x.z); stf<TD>(d + 4 * i + 3, x.w);
    }
}

// ---------------------------------------------------------------------------------------------
// concat / repeat / pad / upscale / timestep embedding / get_rows / misc
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_concat(b200_td a, b200_td b, b200_td d, int dim, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t idx[4];
        int64_t r = i;
        idx[0] = r % d.ne[0]; r /= d.ne[0];
        idx[1] = r % d.ne[1]; r /= d.ne[1];
        idx[2] = r % d.ne[2]; idx[3] = r / d.ne[2];
        char* pd = (char*)d.data + idx[0] * d.nb[0] + idx[1] * d.nb[1] + idx[2] * d.nb[2] + idx[3] * d.nb[3];
        const b200_td* srcp = &a;
        if (idx[dim] >= a.ne[dim]) { idx[dim] -= a.ne[dim]; srcp = &b; }
        const char* ps = (const char*)srcp->data + idx[0] * srcp->nb[0] + idx[1] * srcp->nb[1] + idx[2] * srcp->nb[2] + idx[3] * srcp->nb[3];
        *(T*)pd = *(const T*)ps;
    }
}

template <typename T>
__global__ void k_repeat(b200_td a, b200_td d, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        int64_t i0 = r % d.ne[0]; r /= d.ne[0];
        int64_t i1 = r % d.ne[1]; r /= d.ne[1];
        int64_t i2 = r % d.ne[2]; int64_t i3 = r / d.ne[2];
        *(T*)((char*)d.data + i0 * d.nb[0] + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]) =
            *(const T*)((const char*)a.data + (i0 % a.ne[0]) * a.nb[0] + (i1 % a.ne[1]) * a.nb[1] + (i2 % a.ne[2]) * a.nb[2] + (i3 % a.ne[3]) * a.nb[3]);
    }
}

struct PadParams { int32_t lp[4], rp[4]; int circular; };
__global__ void k_pad_f32(b200_td a, b200_td d, PadParams p, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t idx[4];
        int64_t r = i;
        idx[0] = r % d.ne[0]; r /= d.ne[0];
        idx[1] = r % d.ne[1]; r /= d.ne[1];
        idx[2] = r % d.ne[2]; idx[3] = r / d.ne[2];
        float v = 0.f;
        bool inside = true;
        int64_t s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s[k] = idx[k] - p.lp[k];
            if (p.circular) {
                s[k] = ((s[k] % a.ne[k]) + a.ne[k]) % a.ne[k];
            } else if (s[k] < 0 || s[k] >= a.ne[k]) {
                inside = false;
            }
        }
        if (inside) v = *(const float*)((const char*)a.data + s[0] * a.nb[0] + s[1] * a.nb[1] + s[2] * a.nb[2] + s[3] * a.nb[3]);
        *(float*)((char*)d.data + idx[0] * d.nb[0] + idx[1] * d.nb[1] + idx[2] * d.nb[2] + idx[3] * d.nb[3]) = v;
    }
}

// mode 0 nearest, 1 bilinear (align_corners flag folded into sf/pixel_offset by the host like ops.cpp:7848-7860)
__global__ void k_upscale_f32(b200_td a, b200_td d, int mode, float sf0, float sf1, float sf2, float sf3, float pixel_offset, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        int64_t i0 = r % d.ne[0]; r /= d.ne[0];
        int64_t i1 = r % d.ne[1]; r /= d.ne[1];
        int64_t i2 = r % d.ne[2]; int64_t i3 = r / d.ne[2];
        int64_t s2 = (int64_t)(i2 / sf2), s3 = (int64_t)(i3 / sf3);
        const char* base = (const char*)a.data + s2 * a.nb[2] + s3 * a.nb[3];
        float v;
        if (mode == 0) {
            int64_t s0 = (int64_t)(i0 / sf0), s1 = (int64_t)(i1 / sf1);
            v = *(const float*)(base + s0 * a.nb[0] + s1 * a.nb[1]);
        } else {
            float y = ((float)i1 + pixel_offset) / sf1 - pixel_offset;
            int64_t y0 = (int64_t)floorf(y), y1 = y0 + 1;
            y0 = max((int64_t)0, min(y0, a.ne[1] - 1));
            y1 = max((int64_t)0, min(y1, a.ne[1] - 1));
            float dy = fmaxf(0.f, fminf(y - (float)y0, 1.f));
            float x = ((float)i0 + pixel_offset) / sf0 - pixel_offset;
            int64_t x0 = (int64_t)floorf(x), x1 = x0 + 1;
            x0 = max((int64_t)0, min(x0, a.ne[0] - 1));
            x1 = max((int64_t)0, min(x1, a.ne[0] - 1));
            float dx = fmaxf(0.f, fminf(x - (float)x0, 1.f));
            float va = *(const float*)(base + x0 * a.nb[0] + y0 * a.nb[1]);
            float vb = *(const float*)(base + x1 * a.nb[0] + y0 * a.nb[1]);
            float vc = *(const float*)(base + x0 * a.nb[0] + y1 * a.nb[1]);
            float vd = *(const float*)(base + x1 * a.nb[0] + y1 * a.nb[1]);
            v = va * (1 - dx) * (1 - dy) + vb * dx * (1 - dy) + vc * (1 - dx) * dy + vd * dx * dy;
        }
        *(float*)((char*)d.data + i0 * d.nb[0] + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]) = v;
    }
}

__global__ void k_timestep_embedding(const float* __restrict__ t, char* dst, int64_t nb1, int dim, int max_period, int64_t n_t) {
    pdl_wait();
    pdl_launch_dependents();
    int half = dim / 2;
    int64_t i = blockIdx.y;
    if (i >= n_t) return;
    float* out = (float*)(dst + i * nb1);
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < half) {
        float timestep = t[i];
        float freq = expf(-logf((float)max_period) * j / half);
        float arg = timestep * freq;
        out[j] = cosf(arg);
        out[j + half] = sinf(arg);
    }
    if ((dim & 1) && j == 0) out[2 * half] = 0.f;
}

template <typename TS>
__global__ void k_get_rows(b200_td src, b200_td idx, b200_td d, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    // dst[i0, r, b2, b3] = src[i0, idx[r, b2, b3], b2, b3]
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        int64_t i0 = r % d.ne[0]; r /= d.ne[0];
        int64_t i1 = r % d.ne[1]; r /= d.ne[1];
        int64_t i2 = r % d.ne[2]; int64_t i3 = r / d.ne[2];
        int32_t row = *(const int32_t*)((const char*)idx.data + i1 * idx.nb[0] + i2 * idx.nb[1] + i3 * idx.nb[2]);
        float v = ldf<TS>((const char*)src.data + i0 * src.nb[0] + (int64_t)row * src.nb[1] + i2 * src.nb[2] + i3 * src.nb[3]);
        *(float*)((char*)d.data + i0 * d.nb[0] + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]) = v;
    }
}

__global__ void k_arange(float* d, int64_t n, float start, float step) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = start + step * (float)i;
}
__global__ void k_fill(float* d, int64_t n, float v) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = v;
}

// GLU family: dst[i0] = act(a[i0]) * b[i0]   (a/b are the halves of one tensor when b == nullptr on the host side)
__global__ void k_glu_f32(const char* a, const char* b, char* d, int64_t nc, int64_t nrows, int64_t a_nb1, int64_t b_nb1, int64_t d_nb1, int glu_op) {
    pdl_wait();
    pdl_launch_dependents();
    int64_t n = nc * nrows;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i % nc, r = i / nc;
        float x = *(const float*)(a + r * a_nb1 + c * 4);
        float g = *(const float*)(b + r * b_nb1 + c * 4);
        float y;
        switch (glu_op) {
            case GGML_GLU_OP_REGLU: y = fmaxf(x, 0.f); break;
            case GGML_GLU_OP_GEGLU: y = apply_unary(GGML_UNARY_OP_GELU, x, 0, 0); break;
            case GGML_GLU_OP_SWIGLU: y = apply_unary(GGML_UNARY_OP_SILU, x, 0, 0); break;
            case GGML_GLU_OP_GEGLU_ERF: y = apply_unary(GGML_UNARY_OP_GELU_ERF, x, 0, 0); break;
            case GGML_GLU_OP_GEGLU_QUICK: y = apply_unary(GGML_UNARY_OP_GELU_QUICK, x, 0, 0); break;
            default: y = x;
        }
        *(float*)(d + r * d_nb1 + c * 4) = y * g;
    }
}

// one warp per row
__global__ void k_sum_rows(b200_td a, b200_td d, int64_t nrows, bool mean) {
    pdl_wait();
    pdl_launch_dependents();
    int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= nrows) return;
    int lane = threadIdx.x & 31;
    int64_t i1 = row % a.ne[1], r = row / a.ne[1];
    int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
    const char* p = (const char*)a.data + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3];
    float s = 0.f;
    for (int64_t i = lane; i < a.ne[0]; i += 32) s += *(const float*)(p + i * a.nb[0]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) *(float*)((char*)d.data + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]) = mean ? s / (float)a.ne[0] : s;
}

// pack rows: src logical [K, R1, R2, R3] (any strides, unit stride along K not required) -> dense [R][kpad] of TD, zero padded
template <typename TS, typename TD>
__global__ void k_pack_rows(b200_td a, TD* __restrict__ d, int64_t kpad, int64_t nrows) {
    pdl_wait();
    pdl_launch_dependents();
    int64_t n = nrows * kpad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = i % kpad, row = i / kpad;
        float v = 0.f;
        if (k < a.ne[0]) {
            int64_t i1 = row % a.ne[1], r = row / a.ne[1];
            int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
            v = ldf<TS>((const char*)a.data + k * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
        }
        stf<TD>(d + i, v);
    }
}


// Activation operand of a Q8_0-weight contraction.  The CPU oracle quantises every activation row to Q8_0 blocks before the dot product
// (ggml-cpu.c:1480-1510 -> quantize_row_q8_0: per 32 values d = max|x| / 127 stored as f16, q = round(x / d)) and multiplies int8 x int8
// with the two block scales.  This kernel applies the SAME quantisation and writes d * q back as f16 (exact in f32, one f16 rounding):
// the tensor-core GEMM then contracts the very values the oracle contracts, instead of the un-quantised activations.
// One warp per (row, 32-value block); dense [rows][kpad] f16 out, kpad % 32 == 0, K zero padded.
__global__ void k_pack_rows_q8_roundtrip(b200_td a, __half* __restrict__ d, int64_t kpad, int64_t nrows) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t nblk = kpad / 32;
    const int64_t total = nrows * nblk;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t row = w / nblk, blk = w - row * nblk;
        const int64_t k = blk * 32 + lane;
        float v = 0.f;
        if (k < a.ne[0]) {
            int64_t i1 = row % a.ne[1], r = row / a.ne[1];
            int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
            v = *(const float*)((const char*)a.data + k * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
        }
        float amax = fabsf(v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const float dq = amax / 127.0f;
        const float id = dq != 0.f ? 1.0f / dq : 0.f;
        const float dh = __half2float(__float2half_rn(dq));       // the scale as stored (f16) and used by the oracle's dot product
        const float q = rintf(v * id);                             // round to nearest even, like the AVX quantiser the oracle runs
        d[row * kpad + k] = __float2half_rn(dh * q);
    }
}

// f32 operand -> (hi, lo) pair for the 3xTF32 contraction: hi = x with the 13 low mantissa bits cleared (exactly representable in TF32,
// whatever rounding the tensor core applies to its inputs), lo = x - hi (exact in f32).  Dense [rows][kpad] f32 each, K zero padded.
template <typename TS>
__global__ void k_split_tf32(b200_td a, float* __restrict__ hi, float* __restrict__ lo, int64_t kpad, int64_t nrows) {
    pdl_wait();
    pdl_launch_dependents();
    int64_t n = nrows * kpad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = i % kpad, row = i / kpad;
        float v = 0.f;
        if (k < a.ne[0]) {
            int64_t i1 = row % a.ne[1], r = row / a.ne[1];
            int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
            v = ldf<TS>((const char*)a.data + k * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
        }
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}

// 16-bit tiled transpose through shared memory: out[z][r = d][c = l] = in[z][l][d]
__global__ void k_transpose_16(const char* __restrict__ src, uint16_t* __restrict__ dst, int D, int L, int64_t Lpad, int64_t nb1, int64_t nb2, int64_t nb3,
                               int ne2) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ uint16_t tile[32][34];
    const int z = blockIdx.z;
    const int i2 = z % ne2, i3 = z / ne2;
    const char* sb = src + i2 * nb2 + i3 * nb3;
    uint16_t* db = dst + (int64_t)z * D * Lpad;
    const int l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {          // rows of the source tile: l; columns: d (contiguous)
        const int l = l0 + k, d = d0 + threadIdx.x;
        if (l < L && d < D) tile[k][threadIdx.x] = *(const uint16_t*)(sb + (int64_t)l * nb1 + (int64_t)d * 2);
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {          // rows of the destination tile: d; columns: l (contiguous)
        const int d = d0 + k, l = l0 + threadIdx.x;
        if (l < L && d < D) db[(int64_t)d * Lpad + l] = tile[threadIdx.x][k];
    }
}

}  // namespace

// ================================================================================================
// launchers
// ================================================================================================
int b200_launch_geglu(cudaStream_t s, const b200_td& x, const b200_td& gate, const b200_td& dst, void* dst16) {
    if (x.type != GGML_TYPE_F32 || gate.type != GGML_TYPE_F32 || dst.type != GGML_TYPE_F32) return -1;
    if (!same_shape(x, dst) || !same_shape(gate, dst) || !rows4_ok(x) || !rows4_ok(gate) || !rows4_ok(dst)) return -1;
    if (dst16 && (!td_contiguous(dst, 4) || ((uintptr_t)dst16 & 7))) return -1;
    rows4 r;
    if (!make_rows4(dst, &r)) return -1;
    if (r.total == 0) return 0;
    b200_launch(k_geglu_rows4, dim3(grid_for(r.total)), dim3(kThreads), 0, s, (const char*)x.data, st3(x), (const char*)gate.data, st3(gate),
                (char*)dst.data, st3(dst), (__half*)dst16, r);
    return 1;
}

int b200_launch_binary(cudaStream_t s, int op, const b200_td& a, const b200_td& b, const b200_td& dst) {
    switch (op) {
        case B200_ADD: return launch_binary_op<B200_ADD>(s, a, b, dst);
        case B200_SUB: return launch_binary_op<B200_SUB>(s, a, b, dst);
        case B200_MUL: return launch_binary_op<B200_MUL>(s, a, b, dst);
        case B200_DIV: return launch_binary_op<B200_DIV>(s, a, b, dst);
    }
    return -1;
}

int b200_launch_unary(cudaStream_t s, int unary_op, const b200_td& src, const b200_td& dst) {
    return launch_unary_impl(s, src, dst, UnaryParams{unary_op, 0.f, 0.f});
}

int b200_launch_scalar_op(cudaStream_t s, int op, const b200_td& src, const b200_td& dst, float p0, float p1) {
    return launch_unary_impl(s, src, dst, UnaryParams{100 + op, p0, p1});
}

int b200_launch_copy(cudaStream_t s, const b200_td& a, const b200_td& d) {
    int64_t n = td_nelements(d);
    if (n == 0) return 0;
    int64_t es = type_size(a.type), ed = type_size(d.type);
    if (es == 0 || ed == 0) return -1;
    bool ca = td_contiguous(a, es), cd = td_contiguous(d, ed);
    if (a.type == d.type && ca && cd) {
        int64_t bytes = n * es;
        if (bytes % 16 == 0 && (uintptr_t)a.data % 16 == 0 && (uintptr_t)d.data % 16 == 0) {
            b200_launch(k_copy_contig16, dim3(grid_for(bytes / 16)), dim3(kThreads), 0, s, (const uint4*)a.data, (uint4*)d.data, bytes / 16);
        } else {
            cudaMemcpyAsync(d.data, a.data, bytes, cudaMemcpyDeviceToDevice, s);
        }
        return 1;
    }
    if (a.type == GGML_TYPE_F32 && ca && cd && n % 4 == 0 && (uintptr_t)a.data % 16 == 0 && (d.type == GGML_TYPE_F16 || d.type == GGML_TYPE_BF16)) {
        if (d.type == GGML_TYPE_F16) b200_launch(k_cvt_f32_contig<__half>, dim3(grid_for(n / 4)), dim3(kThreads), 0, s, (const float4*)a.data, (__half*)d.data, n / 4);
        else b200_launch(k_cvt_f32_contig<__nv_bfloat16>, dim3(grid_for(n / 4)), dim3(kThreads), 0, s, (const float4*)a.data, (__nv_bfloat16*)d.data, n / 4);
        return 1;
    }
    // f32 -> f32 transposing copy: dst contiguous, same shape, src unit stride on another axis
    if (a.type == GGML_TYPE_F32 && d.type == GGML_TYPE_F32 && cd && a.nb[0] != 4 && a.ne[0] == d.ne[0] && a.ne[1] == d.ne[1] &&
        a.ne[2] == d.ne[2] && a.ne[3] == d.ne[3]) {
        int ax = -1;
        for (int k = 1; k < 4; ++k)
            if (a.nb[k] == 4 && a.ne[k] > 1) { ax = k; break; }
        if (ax > 0 && a.ne[0] >= 8 && a.ne[ax] >= 8) {
            int o1 = -1, o2 = -1;
            for (int k = 1; k < 4; ++k)
                if (k != ax) { if (o1 < 0) o1 = k; else o2 = k; }
            int64_t nz = d.ne[o1] * d.ne[o2];
            if (nz <= 65535 && (d.ne[ax] + 31) / 32 <= 65535) {
                dim3 grid((unsigned)((d.ne[0] + 31) / 32), (unsigned)((d.ne[ax] + 31) / 32), (unsigned)nz);
                b200_launch(k_transpose_f32, dim3(grid), dim3(dim3(32, 8)), 0, s, a, d, ax, o1, o2);
                return 1;
            }
        }
    }
    if (a.type == d.type && es == 4 && same_shape(a, d) && rows4_ok(a) && rows4_ok(d)) {
        rows4 r;
        if (make_rows4(d, &r)) {
            b200_launch(k_copy_rows4, dim3(grid_for(r.total)), dim3(kThreads), 0, s, (const char*)a.data, st3(a), (char*)d.data, st3(d), r);
            return 1;
        }
    }
    unsigned g = grid_for(n);
#define CP(TS, TD) b200_launch(k_copy_generic<TS, TD>, dim3(g), dim3(kThreads), 0, s, a, d, n)
    if (a.type == d.type) {
        if (es == 4) b200_launch(k_copy_raw<uint32_t>, dim3(g), dim3(kThreads), 0, s, a, d, n);
        else if (es == 2) b200_launch(k_copy_raw<uint16_t>, dim3(g), dim3(kThreads), 0, s, a, d, n);
        else if (es == 1) b200_launch(k_copy_raw<uint8_t>, dim3(g), dim3(kThreads), 0, s, a, d, n);
        else b200_launch(k_copy_raw<uint64_t>, dim3(g), dim3(kThreads), 0, s, a, d, n);
    } else if (a.type == GGML_TYPE_F32 && d.type == GGML_TYPE_F16) CP(float, __half);
    else if (a.type == GGML_TYPE_F32 && d.type == GGML_TYPE_BF16) CP(float, __nv_bfloat16);
    else if (a.type == GGML_TYPE_F16 && d.type == GGML_TYPE_F32) CP(__half, float);
    else if (a.type == GGML_TYPE_BF16 && d.type == GGML_TYPE_F32) CP(__nv_bfloat16, float);
    else if (a.type == GGML_TYPE_F16 && d.type == GGML_TYPE_BF16) CP(__half, __nv_bfloat16);
    else if (a.type == GGML_TYPE_BF16 && d.type == GGML_TYPE_F16) CP(__nv_bfloat16, __half);
    else return -1;
#undef CP
    return 1;
}

int b200_launch_concat(cudaStream_t s, const b200_td& a, const b200_td& b, const b200_td& d, int dim) {
    int64_t n = td_nelements(d);
    if (n == 0) return 0;
    int64_t es = type_size(d.type);
    if (es == 4 && dim >= 1 && rows4_ok(a) && rows4_ok(b) && rows4_ok(d) && a.ne[dim] < (1ll << 31)) {
        rows4 r;
        if (make_rows4(d, &r)) {
            b200_launch(k_concat_rows4, dim3(grid_for(r.total)), dim3(kThreads), 0, s, (const char*)a.data, st3(a), (const char*)b.data, st3(b), (char*)d.data, st3(d), r, dim,
                                                                 (uint32_t)a.ne[dim]);
            return 1;
        }
    }
    if (es == 4 && dim == 0 && a.ne[0] % 4 == 0 && b.ne[0] % 4 == 0 && rows4_ok(a) && rows4_ok(b) && rows4_ok(d) && a.ne[0] / 4 < (1ll << 31)) {
        rows4 r;
        if (make_rows4(d, &r)) {
            b200_launch(k_concat_dim0_rows4, dim3(grid_for(r.total)), dim3(kThreads), 0, s, (const char*)a.data, st3(a), (const char*)b.data, st3(b), (char*)d.data,
                        st3(d), r, (uint32_t)(a.ne[0] / 4));
            return 1;
        }
    }
    if (es == 4) b200_launch(k_concat<uint32_t>, dim3(grid_for(n)), dim3(kThreads), 0, s, a, b, d, dim, n);
    else if (es == 2) b200_launch(k_concat<uint16_t>, dim3(grid_for(n)), dim3(kThreads), 0, s, a, b, d, dim, n);
    else return -1;
    return 1;
}

int b200_launch_repeat(cudaStream_t s, const b200_td& a, const b200_td& d) {
    int64_t n = td_nelements(d);
    if (n == 0) return 0;
    int64_t es = type_size(d.type);
    if (es == 4) b200_launch(k_repeat<uint32_t>, dim3(grid_for(n)), dim3(kThreads), 0, s, a, d, n);
    else if (es == 2) b200_launch(k_repeat<uint16_t>, dim3(grid_for(n)), dim3(kThreads), 0, s, a, d, n);
    else return -1;
    return 1;
}

int b200_launch_pad(cudaStream_t s, const b200_td& a, const b200_td& d, const int32_t* pads, bool circular) {
    int64_t n = td_nelements(d);
    if (n == 0) return 0;
    PadParams p;
    for (int k = 0; k < 4; ++k) { p.lp[k] = pads[2 * k]; p.rp[k] = pads[2 * k + 1]; }
    p.circular = circular;
    b200_launch(k_pad_f32, dim3(grid_for(n)), dim3(kThreads), 0, s, a, d, p, n);
    return 1;
}

int b200_launch_upscale(cudaStream_t s, const b200_td& a, const b200_td& d, int mode_flags) {
    int64_t n = td_nelements(d);
    if (n == 0) return 0;
    float sf0 = (float)d.ne[0] / a.ne[0], sf1 = (float)d.ne[1] / a.ne[1], sf2 = (float)d.ne[2] / a.ne[2], sf3 = (float)d.ne[3] / a.ne[3];
    float pixel_offset = 0.5f;
    int mode = mode_flags & 0xFF;
    if (mode_flags & GGML_SCALE_FLAG_ALIGN_CORNERS) {
        pixel_offset = 0.f;
        sf0 = d.ne[0] > 1 && a.ne[0] > 1 ? (float)(d.ne[0] - 1) / (a.ne[0] - 1) : sf0;
        sf1 = d.ne[1] > 1 && a.ne[1] > 1 ? (float)(d.ne[1] - 1) / (a.ne[1] - 1) : sf1;
    }
    int m = mode == GGML_SCALE_MODE_NEAREST ? 0 : 1;
    b200_launch(k_upscale_f32, dim3(grid_for(n)), dim3(kThreads), 0, s, a, d, m, sf0, sf1, sf2, sf3, pixel_offset, n);
    return 1;
}

int b200_launch_timestep_embedding(cudaStream_t s, const b200_td& src, const b200_td& d, int dim, int max_period) {
    int half = dim / 2;
    dim3 grid((unsigned)((std::max(half, 1) + 127) / 128), (unsigned)src.ne[0]);
    b200_launch(k_timestep_embedding, dim3(grid), dim3(128), 0, s, (const float*)src.data, (char*)d.data, d.nb[1], dim, max_period, src.ne[0]);
    return 1;
}

int b200_launch_get_rows(cudaStream_t s, const b200_td& src, const b200_td& idx, const b200_td& d) {
    int64_t n = td_nelements(d);
    if (n == 0) return 0;
    if (src.type == GGML_TYPE_F32) b200_launch(k_get_rows<float>, dim3(grid_for(n)), dim3(kThreads), 0, s, src, idx, d, n);
    else if (src.type == GGML_TYPE_F16) b200_launch(k_get_rows<__half>, dim3(grid_for(n)), dim3(kThreads), 0, s, src, idx, d, n);
    else if (src.type == GGML_TYPE_BF16) b200_launch(k_get_rows<__nv_bfloat16>, dim3(grid_for(n)), dim3(kThreads), 0, s, src, idx, d, n);
    else return -1;
    return 1;
}

int b200_launch_arange(cudaStream_t s, const b200_td& d, float start, float step) {
    int64_t n = td_nelements(d);
    b200_launch(k_arange, dim3(grid_for(n)), dim3(kThreads), 0, s, (float*)d.data, n, start, step);
    return 1;
}

int b200_launch_fill(cudaStream_t s, const b200_td& d, float v) {
    int64_t n = td_nelements(d);
    b200_launch(k_fill, dim3(grid_for(n)), dim3(kThreads), 0, s, (float*)d.data, n, v);
    return 1;
}

int b200_launch_glu(cudaStream_t s, int glu_op, const b200_td& a, const b200_td* b, const b200_td& d, bool swapped) {
    int64_t nc = d.ne[0], nrows = d.ne[1] * d.ne[2] * d.ne[3];
    if (nc * nrows == 0) return 0;
    const char* pa = (const char*)a.data;
    const char* pb;
    int64_t b_nb1;
    if (b) {
        pb = (const char*)b->data;
        b_nb1 = b->nb[1];
    } else {
        // single tensor: first half = x, second half = gate (swapped flips)
        pb = pa + (swapped ? 0 : nc * 4);
        pa = pa + (swapped ? nc * 4 : 0);
        b_nb1 = a.nb[1];
    }
    b200_launch(k_glu_f32, dim3(grid_for(nc * nrows)), dim3(kThreads), 0, s, pa, pb, (char*)d.data, nc, nrows, a.nb[1], b_nb1, d.nb[1], glu_op);
    return 1;
}

int b200_launch_sum_rows(cudaStream_t s, const b200_td& a, const b200_td& d, bool mean) {
    int64_t nrows = a.ne[1] * a.ne[2] * a.ne[3];
    if (nrows == 0) return 0;
    b200_launch(k_sum_rows, dim3((unsigned)((nrows + 7) / 8)), dim3(256), 0, s, a, d, nrows, mean);
    return 1;
}

int b200_launch_pack_rows(cudaStream_t s, const b200_td& a, void* dst, int dst_type, int64_t kpad) {
    int64_t nrows = a.ne[1] * a.ne[2] * a.ne[3];
    int64_t n = nrows * kpad;
    if (n == 0) return 0;
    if (a.type == GGML_TYPE_F32 && (dst_type == GGML_TYPE_F16 || dst_type == GGML_TYPE_BF16) && a.nb[0] == 4 && kpad % 8 == 0 &&
        (uintptr_t)a.data % 16 == 0 && a.nb[1] % 16 == 0 && a.nb[2] % 16 == 0 && a.nb[3] % 16 == 0 && (uintptr_t)dst % 16 == 0 && a.ne[0] < (1ll << 31)) {
        b200_td shape = a;
        shape.ne[0] = kpad / 2;      // make_rows4 divides by 4: cpr = kpad / 8
        rows4 r;
        if (make_rows4(shape, &r)) {
            if (dst_type == GGML_TYPE_F16) b200_launch(k_pack_rows8<__half>, dim3(grid_for(r.total)), dim3(kThreads), 0, s, (const char*)a.data, st3(a), (__half*)dst, r, (uint32_t)a.ne[0]);
            else b200_launch(k_pack_rows8<__nv_bfloat16>, dim3(grid_for(r.total)), dim3(kThreads), 0, s, (const char*)a.data, st3(a), (__nv_bfloat16*)dst, r, (uint32_t)a.ne[0]);
            return 1;
        }
    }
    unsigned g = grid_for(n);
#define PK(TS, TD) b200_launch(k_pack_rows<TS, TD>, dim3(g), dim3(kThreads), 0, s, a, (TD*)dst, kpad, nrows)
    if (a.type == GGML_TYPE_F32 && dst_type == GGML_TYPE_F16) PK(float, __half);
    else if (a.type == GGML_TYPE_F32 && dst_type == GGML_TYPE_BF16) PK(float, __nv_bfloat16);
    else if (a.type == GGML_TYPE_F32 && dst_type == GGML_TYPE_F32) PK(float, float);
    else if (a.type == GGML_TYPE_F16 && dst_type == GGML_TYPE_F16) PK(__half, __half);
    else if (a.type == GGML_TYPE_BF16 && dst_type == GGML_TYPE_BF16) PK(__nv_bfloat16, __nv_bfloat16);
    else if (a.type == GGML_TYPE_F16 && dst_type == GGML_TYPE_F32) PK(__half, float);
    else if (a.type == GGML_TYPE_BF16 && dst_type == GGML_TYPE_F32) PK(__nv_bfloat16, float);
    else if (a.type == GGML_TYPE_F16 && dst_type == GGML_TYPE_BF16) PK(__half, __nv_bfloat16);
    else if (a.type == GGML_TYPE_BF16 && dst_type == GGML_TYPE_F16) PK(__nv_bfloat16, __half);
    else return -1;
#undef PK
    return 1;
}


int b200_launch_pack_rows_q8_roundtrip(cudaStream_t s, const b200_td& a, void* dst_f16, int64_t kpad) {
    const int64_t nrows = a.ne[1] * a.ne[2] * a.ne[3];
    if (nrows * kpad == 0) return 0;
    if (a.type != GGML_TYPE_F32 || kpad % 32 != 0) return -1;
    b200_launch(k_pack_rows_q8_roundtrip, dim3(grid_for(nrows * kpad)), dim3(kThreads), 0, s, a, (__half*)dst_f16, kpad, nrows);
    return 1;
}

int b200_launch_split_tf32(cudaStream_t s, const b200_td& a, float* hi, float* lo, int64_t kpad) {
    const int64_t nrows = a.ne[1] * a.ne[2] * a.ne[3];
    const int64_t n = nrows * kpad;
    if (n == 0) return 0;
    if (a.type != GGML_TYPE_F32) return -1;
    b200_launch(k_split_tf32<float>, dim3(grid_for(n)), dim3(kThreads), 0, s, a, hi, lo, kpad, nrows);
    return 1;
}

int b200_launch_transpose_f16(cudaStream_t s, const b200_td& src, void* dst, int64_t Lpad) {
    const int64_t D = src.ne[0], L = src.ne[1], Z = src.ne[2] * src.ne[3];
    if (D == 0 || L == 0 || Z == 0) return 0;
    if (src.nb[0] != 2 || Z > 65535 || (D + 31) / 32 > 65535) return -1;
    dim3 grid((unsigned)((L + 31) / 32), (unsigned)((D + 31) / 32), (unsigned)Z);
    b200_launch(k_transpose_16, dim3(grid), dim3(dim3(32, 8)), 0, s, (const char*)src.data, (uint16_t*)dst, (int)D, (int)L, Lpad, src.nb[1], src.nb[2], src.nb[3], (int)src.ne[2]);
    return 1;
}
