"""graph_f64.py -- TEST INFRASTRUCTURE (the arbiter of the whole-model parity tests), never on the product path.

Evaluates a ggml graph exported by the harness (sdh_model_export_graph: the graph the reference's own builder emits for a model call,
with the data of every leaf) in float64 with NO intermediate rounding: f16 / bf16 weights enter with their exact values, activations
are never rounded to the weight type, sums are float64.  That is the mathematical value both backends approximate; the parity tests
ask `|gpu - truth| <= |cpu - truth| (+ margin)` instead of only `|gpu - cpu| < tol`.

Semantics follow the reference's op definitions (file:line in /root/reference/ggml):
  MUL_MAT             src/ggml.c:3282 (dst[n][m] = sum_k a[m][k] * b[n][k], src0 batch broadcast by block repeat)
  IM2COL              src/ggml-cpu/ops.cpp ggml_compute_forward_im2col_f16 (column index ic*KH*KW + kh*KW + kw)
  GROUP_NORM          src/ggml-cpu/ops.cpp:4079-4152     NORM / RMS_NORM  ops.cpp ggml_compute_forward_norm_f32 / rms_norm_f32
  SOFT_MAX            ops.cpp ggml_compute_forward_soft_max_f32 (scale, additive mask)
  UPSCALE (nearest)   ops.cpp ggml_compute_forward_upscale_f32         CONCAT  ops.cpp ggml_compute_forward_concat
  TIMESTEP_EMBEDDING  ops.cpp:8278-8309                                 FLASH_ATTN_EXT  src/ggml.c:5476 (layout), ops.cpp:8468
  GELU is the tanh form (ggml-cpu/vec.h ggml_gelu_f32 without its f16 lookup table), SILU x * sigmoid(x).
View ops (RESHAPE / VIEW / PERMUTE / TRANSPOSE) and in-place results are modelled exactly like ggml does: every tensor is a strided
window (ne, nb, view offset) onto the flat buffer of its view root.
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import numpy as np
from numpy.lib.stride_tricks import as_strided

ES = {"f32": 4, "f16": 2, "bf16": 2, "i32": 4, "q8_0": None}
VIEW_OPS = {"RESHAPE", "VIEW", "PERMUTE", "TRANSPOSE", "NONE"}


def _f(i: int) -> float:
    return float(np.array([i], np.int32).view(np.float32)[0])


class Graph:
    def __init__(self, prefix):
        prefix = Path(prefix)
        self.meta = json.loads(Path(str(prefix) + ".json").read_text())
        # memory-mapped, leaves converted to float64 on first use and every buffer dropped after its last reader: the whole-model graphs
        # carry 1-7 GB of leaf data, and freshly faulted pages are what this evaluation costs on a memory-reclaimed VM (measured: 45 MB/s)
        self.blob = np.memmap(str(prefix) + ".bin", dtype=np.float32, mode="r")
        self.T = self.meta["tensors"]
        self.roots = {}    # tensor id -> flat float64 buffer (only for view roots)
        self.root_of = {}  # tensor id -> (root id, element offset)
        self.leaf = {}     # root id -> (offset, count) in the blob, not yet converted

    # ------------------------------------------------------------------ storage model
    def _nelem(self, t):
        n = 1
        for v in t["ne"]:
            n *= v
        return n

    def _bind(self, t):
        i = t["id"]
        es = ES[t["type"]]
        if t["view_src"] >= 0:
            r, off = self.root_of[t["view_src"]]
            res = ES[self.T[r]["type"]] or 4
            self.root_of[i] = (r, off + t["view_offs"] // res)
            return
        n = self._nelem(t)
        if t["op"] == "NONE" and t["data"] >= 0:
            self.leaf[i] = (t["data"], n)
        else:
            self.leaf[i] = (-1, n)
        self.root_of[i] = (i, 0)

    def _root(self, r):
        buf = self.roots.get(r)
        if buf is None:
            off, n = self.leaf[r]
            buf = np.asarray(self.blob[off:off + n], dtype=np.float64) if off >= 0 else np.zeros(n, np.float64)
            self.roots[r] = buf
        return buf

    def arr(self, i):
        """numpy view of tensor i, numpy axis order = reversed ggml order (ne[3], ne[2], ne[1], ne[0])."""
        t = self.T[i]
        r, off = self.root_of[i]
        es = ES[t["type"]]
        root = self._root(r)
        if es is None:     # quantised leaf: exported dequantised and contiguous
            return root.reshape(t["ne"][::-1])
        strides = tuple(int(b // es) * 8 for b in t["nb"][::-1])
        return as_strided(root[off:], shape=tuple(t["ne"][::-1]), strides=strides, writeable=True)

    # ------------------------------------------------------------------ ops
    def run(self, progress=False):
        for t in self.T:
            self._bind(t)
        order = self.meta["nodes"]
        # last working node that touches each root (as a source or as its destination, through any chain of views)
        last = {}
        for k, i in enumerate(order):
            t = self.T[i]
            if t["op"] in VIEW_OPS:
                continue
            for j in [i] + [s for s in t["src"] if s >= 0]:
                last[self.root_of[j][0]] = k
        keep = self.root_of[self.meta["result"]][0]
        drop = {}
        for r, k in last.items():
            if r != keep:
                drop.setdefault(k, []).append(r)
        for k, i in enumerate(order):
            t = self.T[i]
            if t["op"] in VIEW_OPS:
                continue
            fn = getattr(self, "op_" + t["op"], None)
            if fn is None:
                raise NotImplementedError(f"graph_f64: op {t['op']} (node {k})")
            out = fn(t, [self.arr(s) if s >= 0 else None for s in t["src"]])
            if out is not None:
                dst = self.arr(i)
                dst[...] = np.asarray(out).reshape(dst.shape)
            out = None
            for r in drop.get(k, ()):
                self.roots.pop(r, None)
            if progress and k % 200 == 0:
                print(f"  f64 node {k}/{len(order)} {t['op']}", flush=True)
        return np.array(self.arr(self.meta["result"]))

    @staticmethod
    def _tile_to(b, shape):
        reps = tuple(s // bs for s, bs in zip(shape, b.shape))
        return b if all(r == 1 for r in reps) else np.tile(b, reps)

    def op_ADD(self, t, s): return s[0] + self._tile_to(s[1], s[0].shape)
    def op_SUB(self, t, s): return s[0] - self._tile_to(s[1], s[0].shape)
    def op_MUL(self, t, s): return s[0] * self._tile_to(s[1], s[0].shape)
    def op_DIV(self, t, s): return s[0] / self._tile_to(s[1], s[0].shape)
    def op_CONT(self, t, s): return np.ascontiguousarray(s[0])
    def op_DUP(self, t, s): return np.ascontiguousarray(s[0])
    def op_SCALE(self, t, s): return s[0] * _f(t["params"][0]) + _f(t["params"][1])
    def op_SQR(self, t, s): return s[0] * s[0]

    def op_CPY(self, t, s):
        return np.ascontiguousarray(s[0]).reshape(-1)      # dst is a view of src[1]; same element count, logical order

    def op_UNARY(self, t, s):
        x = s[0]
        u = t["uop"]
        if u == "SILU":
            return x / (1.0 + np.exp(-x))
        if u == "GELU":
            return 0.5 * x * (1.0 + np.tanh(0.79788456080286535587989211986876 * x * (1.0 + 0.044715 * x * x)))
        if u == "GELU_QUICK":
            return x / (1.0 + np.exp(-1.702 * x))
        if u == "SIGMOID":
            return 1.0 / (1.0 + np.exp(-x))
        if u == "TANH":
            return np.tanh(x)
        if u == "RELU":
            return np.maximum(x, 0.0)
        raise NotImplementedError("unary " + u)

    def op_MUL_MAT(self, t, s):
        a, b = s[0], s[1]            # a (b3, b2, M, K), b (B3, B2, N, K)
        r3, r2 = b.shape[0] // a.shape[0], b.shape[1] // a.shape[1]
        if r3 > 1:
            a = np.repeat(a, r3, axis=0)
        if r2 > 1:
            a = np.repeat(a, r2, axis=1)
        return np.matmul(b, np.swapaxes(a, 2, 3))     # (B3, B2, N, M)

    def op_IM2COL(self, t, s):
        p = t["params"]
        s0, s1, p0, p1, d0, d1, is2d = p[0], p[1], p[2], p[3], p[4], p[5], p[6]
        assert is2d == 1
        kshape = self.T[t["src"][0]]["ne"]          # [KW, KH, IC, OC]
        KW, KH = kshape[0], kshape[1]
        img = s[1]                                   # (N, IC, IH, IW)
        N, IC, IH, IW = img.shape
        OW, OH = t["ne"][1], t["ne"][2]
        pad = np.zeros((N, IC, IH + 2 * p1 + d1 * KH, IW + 2 * p0 + d0 * KW))
        pad[:, :, p1:p1 + IH, p0:p0 + IW] = img
        out = np.zeros((N, OH, OW, IC, KH, KW))
        for kh in range(KH):
            for kw in range(KW):
                patch = pad[:, :, kh * d1: kh * d1 + s1 * OH: s1, kw * d0: kw * d0 + s0 * OW: s0]
                out[:, :, :, :, kh, kw] = patch[:, :, :OH, :OW].transpose(0, 2, 3, 1)
        return out.reshape(N, OH, OW, IC * KH * KW)

    def op_GROUP_NORM(self, t, s):
        x = s[0]                                     # (N, C, H, W)
        ng, eps = t["params"][0], _f(t["params"][1])
        N, C = x.shape[0], x.shape[1]
        cpg = (C + ng - 1) // ng
        y = np.empty_like(x)
        for g in range(ng):
            c0, c1 = g * cpg, min(C, (g + 1) * cpg)
            if c0 >= c1:
                break
            blk = x[:, c0:c1]
            mean = blk.mean(axis=(1, 2, 3), keepdims=True)
            var = ((blk - mean) ** 2).mean(axis=(1, 2, 3), keepdims=True)
            y[:, c0:c1] = (blk - mean) / np.sqrt(var + eps)
        return y

    def op_NORM(self, t, s):
        x = s[0]
        eps = _f(t["params"][0])
        mean = x.mean(axis=-1, keepdims=True)
        var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
        return (x - mean) / np.sqrt(var + eps)

    def op_RMS_NORM(self, t, s):
        x = s[0]
        eps = _f(t["params"][0])
        return x / np.sqrt((x * x).mean(axis=-1, keepdims=True) + eps)

    def op_SOFT_MAX(self, t, s):
        x = s[0] * _f(t["params"][0])
        assert _f(t["params"][1]) == 0.0, "ALiBi not modelled"
        if s[1] is not None:
            m = s[1]
            m = m[..., :x.shape[-2], :]
            x = x + self._tile_to(m, x.shape)
        x = x - x.max(axis=-1, keepdims=True)
        e = np.exp(x)
        return e / e.sum(axis=-1, keepdims=True)

    def op_CONCAT(self, t, s):
        return np.concatenate([s[0], s[1]], axis=3 - t["params"][0])

    def op_UPSCALE(self, t, s):
        x = s[0]
        assert (t["params"][0] & 0xFF) == 0, "only nearest"
        out_shape = tuple(t["ne"][::-1])
        idx = []
        for ax in range(4):
            sf = out_shape[ax] / x.shape[ax]
            idx.append(np.minimum((np.arange(out_shape[ax]) / sf).astype(np.int64), x.shape[ax] - 1))
        return x[np.ix_(*idx)]

    def op_TIMESTEP_EMBEDDING(self, t, s):
        ts = s[0].reshape(-1)
        dim, max_period = t["params"][0], t["params"][1]
        half = dim // 2
        j = np.arange(half, dtype=np.float64)
        freq = np.exp(-math.log(max_period) * j / half)
        arg = ts[:, None] * freq[None, :]
        out = np.zeros((len(ts), t["ne"][0]))
        out[:, :half] = np.cos(arg)
        out[:, half:2 * half] = np.sin(arg)
        return out

    def op_FLASH_ATTN_EXT(self, t, s):
        q, k, v, mask = s[0], s[1], s[2], s[3]       # q (N, H, Lq, d); k (N, Hkv, Lk, d); v (N, Hkv, Lk, dv)
        scale = _f(t["params"][0])
        assert _f(t["params"][1]) == 0.0
        H, Hkv = q.shape[1], k.shape[1]
        if H != Hkv:
            k = np.repeat(k, H // Hkv, axis=1)
            v = np.repeat(v, H // Hkv, axis=1)
        sc = np.matmul(q, np.swapaxes(k, 2, 3)) * scale      # (N, H, Lq, Lk)
        if mask is not None:
            m = mask[..., :q.shape[2], :]
            sc = sc + self._tile_to(m, sc.shape)
        sc = sc - sc.max(axis=-1, keepdims=True)
        e = np.exp(sc)
        p = e / e.sum(axis=-1, keepdims=True)
        o = np.matmul(p, v)                                   # (N, H, Lq, dv)
        return o.transpose(0, 2, 1, 3)                        # ggml dst [dv, H, Lq, N]

    def op_GET_ROWS(self, t, s):
        idx = s[1].astype(np.int64)
        return s[0].reshape(-1, s[0].shape[-1])[idx.reshape(-1)]


def evaluate(prefix, progress=False) -> np.ndarray:
    return Graph(prefix).run(progress)
