"""oracle/ops_ref.py -- TEST INFRASTRUCTURE.  numpy restatement of the reference CPU backend's arithmetic for
the hot-path ops (the ggml ops the UNet / VAE / DiT graphs emit).  Each function cites the reference code it
follows.  Pinned (tests/test_oracle.py) against the reference's own known-answer vectors
(ggml/tests/test-conv2d.cpp:237-360 -> tests/golden/ggml_test_conv2d.json) and against the reference CPU backend
itself (oracle/_ref, compiled from /root/reference) on seeded inputs (tests/golden/cpu_ops_*.npz).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the product path
never does.  Arrays are numpy-ordered, i.e. reversed ggml `ne`: an image is [N, C, H, W], tokens are [N, L, C].
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _f16(x):
    """round-to-nearest-even to IEEE half and back (GGML_CPU_FP32_TO_FP16, ggml-cpu/simd-mappings.h)."""
    return np.asarray(x, F32).astype(np.float16).astype(F32)


def _bf16(x):
    """ggml_compute_fp32_to_bf16 (ggml-impl.h): round-to-nearest-even on the upper 16 bits."""
    u = np.asarray(x, F32).view(np.uint32)
    r = (u + (0x7FFF + ((u >> 16) & 1))) >> 16
    return (r.astype(np.uint32) << 16).view(F32)


# ------------------------------------------------------------------------------------------------
# normalisation   (ggml/src/ggml-cpu/ops.cpp:4079-4152 group_norm; norm_f32; rms_norm_f32)
# ------------------------------------------------------------------------------------------------
def group_norm(x, n_groups: int, eps: float):
    N, C, H, W = x.shape
    cpg = (C + n_groups - 1) // n_groups
    y = np.empty_like(x, dtype=F32)
    for n in range(N):
        for g in range(n_groups):
            c0, c1 = g * cpg, min((g + 1) * cpg, C)
            if c0 >= c1:
                continue
            blk = x[n, c0:c1].astype(np.float64)
            mean = F32(blk.sum() / blk.size)                       # ggml_float (double) sums, float mean
            v = (x[n, c0:c1] - mean).astype(F32)
            var = F32((v.astype(np.float64) ** 2).sum() / blk.size)
            scale = F32(1.0) / np.sqrt(F32(var + F32(eps)))
            y[n, c0:c1] = v * scale
    return y


def norm(x, eps: float):
    mean = x.mean(-1, keepdims=True, dtype=np.float64).astype(F32)
    v = (x - mean).astype(F32)
    var = (v.astype(np.float64) ** 2).mean(-1, keepdims=True).astype(F32)
    return (v * (F32(1.0) / np.sqrt(var + F32(eps)))).astype(F32)


def rms_norm(x, eps: float):
    ms = (x.astype(np.float64) ** 2).mean(-1, keepdims=True).astype(F32)
    return (x * (F32(1.0) / np.sqrt(ms + F32(eps)))).astype(F32)


# ------------------------------------------------------------------------------------------------
# pointwise   (ggml/src/ggml-cpu/vec.h:963-1060)
# ------------------------------------------------------------------------------------------------
def silu(x):
    x = np.asarray(x, F32)
    return (x / (F32(1.0) + np.exp(-x))).astype(F32)


def gelu(x):
    """tanh form, ggml_gelu_f32 (vec.h:968).  NB the default CPU build evaluates it through an f16 lookup table
    (GGML_GELU_FP16, vec.h:988-1001): exact only to ~1e-3 relative."""
    x = np.asarray(x, F32)
    return (F32(0.5) * x * (F32(1.0) + np.tanh(F32(0.79788456080286535587989211986876) * x * (F32(1.0) + F32(0.044715) * x * x)))).astype(F32)


def soft_max(x, mask=None, scale: float = 1.0, max_bias: float = 0.0):
    """ggml_compute_forward_soft_max_f32 (ops.cpp): x*scale + mask, max-subtracted exp, divided by the (double) sum."""
    assert max_bias == 0.0
    w = (x * F32(scale)).astype(F32)
    if mask is not None:
        w = w + np.asarray(mask, F32)
    w = w - w.max(-1, keepdims=True)
    e = np.exp(w).astype(F32)
    return (e / e.sum(-1, keepdims=True, dtype=np.float64)).astype(F32)


def timestep_embedding(t, dim: int, max_period: int = 10000):
    """ops.cpp:8278-8309: [cos(t f_j) | sin(t f_j)], f_j = exp(-ln(max_period) j / half)."""
    t = np.asarray(t, F32).reshape(-1)
    half = dim // 2
    j = np.arange(half, dtype=F32)
    freq = np.exp(-np.log(F32(max_period)) * j / F32(half)).astype(F32)
    arg = t[:, None] * freq[None, :]
    out = np.zeros((t.shape[0], dim), F32)
    out[:, :half] = np.cos(arg)
    out[:, half:2 * half] = np.sin(arg)
    return out


def upscale_nearest(x, factor: int):
    """ops.cpp:7832 (GGML_SCALE_MODE_NEAREST): dst[i] = src[floor(i / sf)]."""
    return np.repeat(np.repeat(x, factor, axis=-1), factor, axis=-2)


# ------------------------------------------------------------------------------------------------
# im2col + mul_mat = conv   (ops.cpp:6426-6500 im2col_f16; ggml-cpu.c:1406 mul_mat; ggml.c:4732 conv_2d)
# ------------------------------------------------------------------------------------------------
def im2col(x, KH, KW, s0=1, s1=1, p0=0, p1=0, d0=1, d1=1, dst_f16=True):
    """x [N, IC, IH, IW] -> [N, OH, OW, IC*KH*KW]; k index = ic*KH*KW + kh*KW + kw; zero outside the image."""
    N, IC, IH, IW = x.shape
    OH = (IH + 2 * p1 - d1 * (KH - 1) - 1) // s1 + 1
    OW = (IW + 2 * p0 - d0 * (KW - 1) - 1) // s0 + 1
    xp = np.zeros((N, IC, IH + 2 * p1, IW + 2 * p0), F32)
    xp[:, :, p1:p1 + IH, p0:p0 + IW] = x
    cols = np.zeros((N, OH, OW, IC, KH, KW), F32)
    for kh in range(KH):
        for kw in range(KW):
            patch = xp[:, :, kh * d1: kh * d1 + (OH - 1) * s1 + 1: s1, kw * d0: kw * d0 + (OW - 1) * s0 + 1: s0]
            cols[:, :, :, :, kh, kw] = patch.transpose(0, 2, 3, 1)
    cols = cols.reshape(N, OH, OW, IC * KH * KW)
    return _f16(cols) if dst_f16 else cols


def mul_mat(w, x, wtype="f32"):
    """dst[n, m] = sum_k w[m, k] x[n, k].  The CPU backend converts x to the weight's vec_dot type first
    (ggml-cpu.c:1430-1513): F16 weights -> x rounded to f16; BF16 -> bf16; F32 -> untouched.  f32 accumulation."""
    w = np.asarray(w, F32)
    x = np.asarray(x, F32)
    if wtype == "f16":
        w, x = _f16(w), _f16(x)
    elif wtype == "bf16":
        w, x = _bf16(w), _bf16(x)
    return (x.astype(np.float64) @ w.astype(np.float64).T).astype(F32)


# ------------------------------------------------------------------------------------------------
# rotary embedding   (Rope::apply_rope, src/model/common/rope.hpp:966-1010, interleaved variant)
# ------------------------------------------------------------------------------------------------
def rope_interleaved(x, pe):
    """x [N, L, H, d] (ggml [d, H, L, N]), pe [L, d/2, 2, 2] = [[cos, -sin], [sin, cos]] -> [N*H, L, d]:
    out[2i + j] = x[2i] * pe[l, i, j, 0] + x[2i + 1] * pe[l, i, j, 1]; the two products are rounded to f32 before the add."""
    x = np.asarray(x, F32)
    pe = np.asarray(pe, F32)
    N, L, H, d = x.shape
    xp = x.transpose(0, 2, 1, 3).reshape(N * H, L, d // 2, 2)
    x0, x1 = xp[..., 0:1], xp[..., 1:2]                                # [NH, L, d/2, 1]
    out = (x0 * pe[None, :, :, :, 0]).astype(F32) + (x1 * pe[None, :, :, :, 1]).astype(F32)
    return out.astype(F32).reshape(N * H, L, d)


# ------------------------------------------------------------------------------------------------
# Q8_0   (block_q8_0 {f16 d; int8 qs[32]}: ggml-common.h:251-255; quantize_row_q8_0_ref / dequantize_row_q8_0: ggml-quants.c;
#         ggml_vec_dot_q8_0_q8_0: ggml-cpu/quants.c -- sum over blocks of int32(sum q_w * q_x) * (d_w * d_x), f32 accumulation)
# ------------------------------------------------------------------------------------------------
def quant_q8_0(x):
    """rows of length K (K % 32 == 0) -> (d [.., K/32] as f16-rounded f32, q [.., K/32, 32] int8).  d = amax / 127 is computed in f32 and the
    quants use 1/d of that f32 value (NOT of the stored f16 d), exactly as quantize_row_q8_0_ref does."""
    x = np.asarray(x, F32)
    b = x.reshape(x.shape[:-1] + (x.shape[-1] // 32, 32))
    amax = np.max(np.abs(b), axis=-1)
    d = (amax / F32(127.0)).astype(F32)
    inv = np.where(d != 0, F32(1.0) / np.where(d != 0, d, F32(1.0)), F32(0.0)).astype(F32)
    v = (b * inv[..., None]).astype(F32)
    q = (np.sign(v) * np.floor(np.abs(v) + F32(0.5))).astype(np.int8)      # roundf: half away from zero
    return _f16(d), q


def dequant_q8_0(dq):
    d, q = dq
    y = (q.astype(F32) * d[..., None]).astype(F32)
    return y.reshape(y.shape[:-2] + (y.shape[-2] * 32,))


def mul_mat_q8_0(w, x):
    """MUL_MAT with Q8_0 weights on the CPU backend: the activation rows are quantised to Q8_0 too (ggml-cpu.c:1480-1510), then
    per 32-block integer dot products scaled by d_w * d_x."""
    dw, qw = quant_q8_0(w)
    dx, qx = quant_q8_0(x)
    sumi = np.einsum("mbk,nbk->nmb", qw.astype(np.int32), qx.astype(np.int32))
    return np.einsum("nmb,mb,nb->nm", sumi.astype(np.float64), dw.astype(np.float64), dx.astype(np.float64)).astype(F32)


def im2col_3d(x, IC, KD, KH, KW, s=(1, 1, 1), p=(0, 0, 0), d=(1, 1, 1), dst_f16=True):
    """GGML_OP_IM2COL_3D (ops.cpp:6625-6709): x [N*IC, ID, IH, IW] -> [N*OD, OH, OW, IC*KD*KH*KW], k = ((ic*KD + kd)*KH + kh)*KW + kw;
    s/p/d are (w, h, d) like the op params."""
    NIC, ID, IH, IW = x.shape
    N = NIC // IC
    OW = (IW + 2 * p[0] - d[0] * (KW - 1) - 1) // s[0] + 1
    OH = (IH + 2 * p[1] - d[1] * (KH - 1) - 1) // s[1] + 1
    OD = (ID + 2 * p[2] - d[2] * (KD - 1) - 1) // s[2] + 1
    xp = np.zeros((N, IC, ID + 2 * p[2], IH + 2 * p[1], IW + 2 * p[0]), F32)
    xp[:, :, p[2]:p[2] + ID, p[1]:p[1] + IH, p[0]:p[0] + IW] = x.reshape(N, IC, ID, IH, IW)
    cols = np.zeros((N, OD, OH, OW, IC, KD, KH, KW), F32)
    for kd in range(KD):
        for kh in range(KH):
            for kw in range(KW):
                patch = xp[:, :, kd * d[2]: kd * d[2] + (OD - 1) * s[2] + 1: s[2], kh * d[1]: kh * d[1] + (OH - 1) * s[1] + 1: s[1],
                           kw * d[0]: kw * d[0] + (OW - 1) * s[0] + 1: s[0]]
                cols[:, :, :, :, :, kd, kh, kw] = patch.transpose(0, 2, 3, 4, 1)
    cols = cols.reshape(N * OD, OH, OW, IC * KD * KH * KW)
    return _f16(cols) if dst_f16 else cols


def conv_3d(w, x, IC, s=(1, 1, 1), p=(0, 0, 0), d=(1, 1, 1)):
    """ggml_conv_3d (ggml.c:4809-4839): im2col_3d(F16) x F16 kernel [OC*IC, KD, KH, KW] -> [N*OC, OD, OH, OW]."""
    OCIC, KD, KH, KW = w.shape
    OC = OCIC // IC
    cols = im2col_3d(x, IC, KD, KH, KW, s, p, d, dst_f16=True)
    NOD, OH, OW, K = cols.shape
    N = x.shape[0] // IC
    OD = NOD // N
    y = mul_mat(w.reshape(OC, K), cols.reshape(NOD * OH * OW, K), "f16")       # [N*OD*OH*OW, OC]
    y = y.reshape(N, OD, OH, OW, OC).transpose(0, 4, 1, 2, 3)
    return np.ascontiguousarray(y.reshape(N * OC, OD, OH, OW), F32)


def conv_2d(w, x, bias=None, s=1, p=0, d=1):
    """ggml_conv_2d (ggml.c:4732-4753): im2col(F16) -> mul_mat against the F16 kernel -> [N, OC, OH, OW] (+ bias)."""
    OC, IC, KH, KW = w.shape
    cols = im2col(x, KH, KW, s, s, p, p, d, d, dst_f16=True)
    N, OH, OW, K = cols.shape
    y = mul_mat(w.reshape(OC, K), cols.reshape(N * OH * OW, K), "f16")          # [N*OH*OW, OC]
    y = y.reshape(N, OH, OW, OC).transpose(0, 3, 1, 2)
    if bias is not None:
        y = y + np.asarray(bias, F32).reshape(1, OC, 1, 1)
    return np.ascontiguousarray(y, F32)


# ------------------------------------------------------------------------------------------------
# attention   (ops.cpp:8468-9176 flash_attn_ext: Q -> F16, K/V F16, f32 softmax)
# ------------------------------------------------------------------------------------------------
def flash_attn_ext(q, k, v, mask=None, scale: float = 1.0):
    """q [N, H, Lq, d] f32, k [N, Hkv, Lk, d] (f16 values), v [N, Hkv, Lk, dv] (f16 values), mask [.., Lq, Lk]
    -> [N, Lq, H, dv] (ggml dst ne = [dv, H, Lq, N])."""
    q, k, v = _f16(q).astype(np.float64), _f16(k).astype(np.float64), _f16(v).astype(np.float64)
    N, H, Lq, d = q.shape
    rk = H // k.shape[1]
    out = np.zeros((N, Lq, H, v.shape[-1]), F32)
    for n in range(N):
        for h in range(H):
            s = (q[n, h] @ k[n, h // rk].T) * scale
            if mask is not None:
                m = np.asarray(mask, F32)
                s = s + _f16(m.reshape(-1, m.shape[-2], m.shape[-1])[0][:Lq])
            s = s - s.max(-1, keepdims=True)
            p = np.exp(s)
            p /= p.sum(-1, keepdims=True)
            out[n, :, h, :] = (p @ v[n, h // rk]).astype(F32)
    return out


def nmse(a, b):
    """test-backend-ops' error metric (ggml/tests/test-backend-ops.cpp:271-285)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300))
