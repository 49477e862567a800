// im2col.cu -- GGML_OP_IM2COL for the unfused graph variant (and test-backend-ops coverage).
// Oracle: ggml/src/ggml-cpu/ops.cpp:6426-6500 (im2col_f16) / im2col_f32: image [IW, IH, IC, N] (f32 or f16) ->
// columns [IC*KH*KW, OW, OH, N] (f16 or f32), k = ic*KH*KW + kh*KW + kw, zero outside the image.
// HBM-bound: algorithmic bytes = |image| read + |columns| written.  In whole-model graphs the fusion pass
// replaces IM2COL+MUL_MAT by an implicit-GEMM kernel that never materialises the columns.
#include "../b200_ops.h"
#include "b200_launch.cuh"

#include <cuda_fp16.h>

namespace {

template <typename TS, typename TD>
__global__ void k_im2col(const char* __restrict__ src, TD* __restrict__ dst, int64_t IW, int64_t IH, int64_t IC, int64_t N, int64_t OW, int64_t OH,
                         int KW, int KH, int s0, int s1, int p0, int p1, int d0, int d1, int64_t nb_row, int64_t nb_ch, int64_t nb_n, int64_t total) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t K = IC * KH * KW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = i % K, r = i / K;
        int64_t ow = r % OW; r /= OW;
        int64_t oh = r % OH; int64_t n = r / OH;
        int64_t ic = k / (KH * KW);
        int kk = (int)(k % (KH * KW));
        int kh = kk / KW, kw = kk % KW;
        int64_t iw = ow * s0 + (int64_t)kw * d0 - p0;
        int64_t ih = oh * s1 + (int64_t)kh * d1 - p1;
        float v = 0.f;
        if (iw >= 0 && iw < IW && ih >= 0 && ih < IH) {
            const char* p = src + n * nb_n + ic * nb_ch + ih * nb_row + iw * (int64_t)sizeof(TS);
            v = (float)(*(const TS*)p);
        }
        dst[i] = (TD)v;
    }
}

// GGML_OP_IM2COL_3D (oracle: ggml-cpu/ops.cpp:6625-6709): image [IW, IH, ID, N*IC] f32 -> columns [IC*KD*KH*KW, OW, OH, N*OD],
// k = ic*KD*KH*KW + kd*KH*KW + kh*KW + kw, zero outside the volume (the Wan patch embedding / causal 3-D convolutions)
template <typename TD>
__global__ void k_im2col_3d(const char* __restrict__ src, TD* __restrict__ dst, int64_t IW, int64_t IH, int64_t ID, int64_t IC, int64_t OW, int64_t OH,
                            int64_t OD, int KW, int KH, int KD, int s0, int s1, int s2, int p0, int p1, int p2, int d0, int d1, int d2, int64_t nb1,
                            int64_t nb2, int64_t nb3, int64_t total) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t KV = (int64_t)KD * KH * KW, K = IC * KV;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = i % K, r = i / K;
        const int64_t ow = r % OW; r /= OW;
        const int64_t oh = r % OH; r /= OH;
        const int64_t od = r % OD;
        const int64_t n = r / OD;
        const int64_t ic = k / KV;
        int kk = (int)(k % KV);
        const int kd = kk / (KH * KW);
        kk -= kd * KH * KW;
        const int kh = kk / KW, kw = kk % KW;
        const int64_t iw = ow * s0 + (int64_t)kw * d0 - p0;
        const int64_t ih = oh * s1 + (int64_t)kh * d1 - p1;
        const int64_t id = od * s2 + (int64_t)kd * d2 - p2;
        float v = 0.f;
        if (iw >= 0 && iw < IW && ih >= 0 && ih < IH && id >= 0 && id < ID)
            v = *(const float*)(src + (n * IC + ic) * nb3 + id * nb2 + ih * nb1 + iw * 4);
        dst[i] = (TD)v;
    }
}

}  // namespace

int b200_launch_im2col_3d(cudaStream_t s, const b200_td& src, const b200_td& dst, int64_t KW, int64_t KH, int64_t KD, int64_t IC, const int32_t* p) {
    if (src.type != GGML_TYPE_F32 || src.nb[0] != 4 || IC <= 0 || src.ne[3] % IC) return -1;
    const int64_t N = src.ne[3] / IC, OD = dst.ne[3] / (N > 0 ? N : 1);
    const int64_t total = dst.ne[0] * dst.ne[1] * dst.ne[2] * dst.ne[3];
    if (total == 0) return 0;
    if (dst.ne[0] != IC * KD * KH * KW) return -1;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffff) blocks = 0x7fffffff;
#define IM3(TD) b200_launch(k_im2col_3d<TD>, dim3((unsigned)blocks), dim3(256), 0, s, (const char*)src.data, (TD*)dst.data, src.ne[0], src.ne[1], src.ne[2], IC, \
                            dst.ne[1], dst.ne[2], OD, (int)KW, (int)KH, (int)KD, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], src.nb[1], src.nb[2],       \
                            src.nb[3], total)
    if (dst.type == GGML_TYPE_F16) IM3(__half);
    else if (dst.type == GGML_TYPE_F32) IM3(float);
    else return -1;
#undef IM3
    return 1;
}

int b200_launch_im2col(cudaStream_t s, const b200_td& src, const b200_td& dst, int64_t KW, int64_t KH, int s0, int s1, int p0, int p1, int d0,
                       int d1, bool is_2d) {
    const int64_t N = is_2d ? src.ne[3] : src.ne[2];
    const int64_t IC = is_2d ? src.ne[2] : src.ne[1];
    const int64_t IH = is_2d ? src.ne[1] : 1;
    const int64_t IW = src.ne[0];
    const int64_t OH = is_2d ? dst.ne[2] : 1;
    const int64_t OW = dst.ne[1];
    if (!is_2d) KH = 1;
    const int64_t nb_n = is_2d ? src.nb[3] : src.nb[2];
    const int64_t nb_ch = is_2d ? src.nb[2] : src.nb[1];
    const int64_t nb_row = is_2d ? src.nb[1] : 0;
    int64_t total = IC * KH * KW * OW * OH * N;
    if (total == 0) return 0;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffff) blocks = 0x7fffffff;
#define IM(TS, TD) b200_launch(k_im2col<TS, TD>, dim3((unsigned)blocks), dim3(256), 0, s, (const char*)src.data, (TD*)dst.data, IW, IH, IC, N, OW, OH, (int)KW, (int)KH, s0, s1, p0, p1, d0, d1, nb_row, nb_ch, nb_n, total)
    if (src.type == GGML_TYPE_F32 && dst.type == GGML_TYPE_F16) IM(float, __half);
    else if (src.type == GGML_TYPE_F32 && dst.type == GGML_TYPE_F32) IM(float, float);
    else if (src.type == GGML_TYPE_F16 && dst.type == GGML_TYPE_F16) IM(__half, __half);
    else if (src.type == GGML_TYPE_F16 && dst.type == GGML_TYPE_F32) IM(__half, float);
    else return -1;
#undef IM
    return 1;
}
