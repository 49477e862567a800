"""ctypes binding of the whole-model harness (include/sd_b200_harness.h).

Host-side plumbing only: loads host/_ref/libsd_harness.so (the reference's unmodified graph
builders + sampler compiled around our synthetic weight manager), registers ggml backend
plugins by path, and exposes numpy in / numpy out calls.  No compute happens in Python.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
HARNESS_SO = REPO / "host" / "_ref" / "libsd_harness.so"
# SDB200_PLUGIN: A/B a differently built plugin (profiling only); the default is the in-tree build
B200_SO = Path(os.environ.get("SDB200_PLUGIN", str(REPO / "stable-diffusion.cpp_b200" / "lib" / "libggml-b200.so")))

FLAG_FLASH_ATTN = 1
FLAG_CONV_DIRECT = 2
FLAG_NO_WEIGHT_VALUES = 4      # place the parameters but do not fill them: graph / claim walks that never run the model


class SdhTensor(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("ne", C.c_int64 * 4)]


def _as_sdh(a: np.ndarray | None):
    """numpy array in ggml order: a.shape == ne[::-1] (numpy slowest-first)."""
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float32)
    t = SdhTensor()
    t.data = a.ctypes.data_as(C.POINTER(C.c_float))
    ne = list(a.shape[::-1]) + [1] * (4 - a.ndim)
    for i in range(4):
        t.ne[i] = ne[i]
    return t, a


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_size_t, C.c_void_p)

STAT_KEYS = ("graphs", "kernel_launches", "nodes_executed", "fused_nodes", "last_graph_ms", "total_graph_ms", "tc_gemm_launches",
             "tc_gemm_flops", "tc_gemm_us", "fused_attn_launches", "cuda_graph_replays", "implicit_convs", "q_read_in_place", "gemv_launches", "rope_launches", "side_stream_launches",
             "gemm_ref_launches", "host_us", "weight_write_graphs", "per_graph_filter_packs", "geglu_epilogues", "cta2_gemm_launches",
             "derived_weight_bytes", "unfused_attention", "peer_exchanges",
             "host_set_us", "host_get_us", "host_compute_us", "host_outside_us", "kv_in_place", "attn_out_f16_only", "gated_residual_epilogues")


class Harness:
    _lib = None

    def __init__(self):
        if Harness._lib is None:
            if not HARNESS_SO.exists():
                raise RuntimeError(f"{HARNESS_SO} missing: run __graft_entry__.build() where /root/reference exists")
            lib = C.CDLL(str(HARNESS_SO), mode=C.RTLD_GLOBAL)
            lib.sdh_load_backend.argtypes = [C.c_char_p]
            lib.sdh_load_backend.restype = C.c_int
            lib.sdh_device_count.restype = C.c_int
            lib.sdh_device_name.argtypes = [C.c_int]
            lib.sdh_device_name.restype = C.c_char_p
            lib.sdh_last_error.restype = C.c_char_p
            lib.sdh_model_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_uint64, C.c_int]
            lib.sdh_model_create.restype = C.c_void_p
            lib.sdh_model_free.argtypes = [C.c_void_p]
            lib.sdh_model_param_bytes.argtypes = [C.c_void_p]
            lib.sdh_model_param_bytes.restype = C.c_size_t
            lib.sdh_model_param_count.argtypes = [C.c_void_p]
            lib.sdh_model_param_count.restype = C.c_int
            P = C.POINTER(SdhTensor)
            lib.sdh_model_out_shape.argtypes = [C.c_void_p, P, C.POINTER(C.c_int64)]
            lib.sdh_model_forward.argtypes = [C.c_void_p, P, P, P, P, P, C.POINTER(C.c_double)]
            lib.sdh_model_dump_graph.argtypes = [C.c_void_p, P, P, P, P, C.c_char_p]
            lib.sdh_model_mailbox_create.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
            lib.sdh_model_mailbox_connect.argtypes = [C.c_void_p, C.c_void_p]
            lib.sdh_model_mailbox_close.argtypes = [C.c_void_p]
            lib.sdh_model_add_to_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_float]
            lib.sdh_model_export_graph.argtypes = [C.c_void_p, P, P, P, P, C.c_char_p]
            lib.sdh_model_last_graph_flops.argtypes = [C.c_void_p]
            lib.sdh_model_last_graph_flops.restype = C.c_double
            lib.sdh_model_last_graph_nodes.argtypes = [C.c_void_p]
            lib.sdh_sample.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_uint64, P, P, P, P, P, P,
                                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)]
            lib.sdh_schedule.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
            lib.sdh_randn.argtypes = [C.c_uint64, C.POINTER(C.c_float), C.c_size_t]
            lib.sdh_run_op.argtypes = [C.c_char_p, C.c_char_p, C.c_int, P, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_float), P, C.c_int]
            lib.sdh_model_backend_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
            lib.sdh_model_set_backend_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
            lib.sdh_sample_split.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_uint64, P, P, P, P, P, P,
                                             C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                             C.c_int, EXCHANGE_FN, C.c_void_p]
            Harness._lib = lib
        self.lib = Harness._lib

    # ------------------------------------------------------------------ registry
    def load_backend(self, so_path) -> int:
        n = self.lib.sdh_load_backend(str(so_path).encode())
        return n

    def devices(self) -> list[str]:
        return [self.lib.sdh_device_name(i).decode() for i in range(self.lib.sdh_device_count())]

    def load_b200(self) -> list[str]:
        """Register this repo's plugin; fails loudly when the CUDA library is missing."""
        if not B200_SO.exists():
            raise RuntimeError(f"{B200_SO} missing: the CUDA backend was not built (no CPU fallback exists)")
        if self.load_backend(B200_SO) < 0:
            raise RuntimeError("libggml-b200.so failed to load: " + self.last_error())
        devs = [d for d in self.devices() if d.startswith("B200_")]
        if not devs:
            raise RuntimeError("libggml-b200.so loaded but registered no device (no usable sm_100 GPU?)")
        return devs

    def last_error(self) -> str:
        return (self.lib.sdh_last_error() or b"").decode()

    def op_name(self, op: int) -> str:
        """ggml_op_name of the reference's ggml (libggml-base.so, already loaded globally by the harness)."""
        fn = self.lib.ggml_op_name
        fn.argtypes = [C.c_int]
        fn.restype = C.c_char_p
        return (fn(int(op)) or b"").decode(errors="replace")

    # ------------------------------------------------------------------ host-side reference math
    def schedule(self, steps: int):
        s = np.zeros(steps + 1, np.float32)
        t = np.zeros(steps, np.float32)
        rc = self.lib.sdh_schedule(steps, s.ctypes.data_as(C.POINTER(C.c_float)), t.ctypes.data_as(C.POINTER(C.c_float)))
        if rc != 0:
            raise RuntimeError(self.last_error())
        return s, t

    def randn(self, seed: int, shape) -> np.ndarray:
        n = int(np.prod(shape))
        a = np.zeros(n, np.float32)
        self.lib.sdh_randn(seed, a.ctypes.data_as(C.POINTER(C.c_float)), n)
        return a.reshape(shape)

    def run_op(self, device: str, op: str, inputs, itypes=None, ip=(), fp=(), n_threads: int = 0) -> np.ndarray:
        """Run one ggml op (sdh_run_op) on `device`.  inputs: list of numpy arrays (numpy order = reversed ggml ne) or None."""
        GGML_TYPES = {"f32": 0, "f16": 1, "bf16": 30, "q8_0": 8}
        n = len(inputs)
        arr = (SdhTensor * 4)()
        keep = []
        for i, a in enumerate(inputs):
            if a is None:
                continue
            s, k = _as_sdh(a)
            arr[i] = s
            keep.append(k)
        it = (C.c_int32 * 4)(*[GGML_TYPES[t] if isinstance(t, str) else int(t) for t in (list(itypes or []) + ["f32"] * 4)[:4]])
        ipa = (C.c_int32 * 16)(*(list(ip) + [0] * 16)[:16])
        fpa = (C.c_float * 8)(*(list(fp) + [0.0] * 8)[:8])
        out = SdhTensor()
        rc = self.lib.sdh_run_op(device.encode(), op.encode(), n, arr, it, ipa, fpa, C.byref(out), n_threads)
        if rc != 0:
            raise RuntimeError(f"run_op({op}) sizing failed: " + self.last_error())
        shape = tuple(out.ne)[::-1]
        res = np.empty(shape, np.float32)
        out.data = res.ctypes.data_as(C.POINTER(C.c_float))
        rc = self.lib.sdh_run_op(device.encode(), op.encode(), n, arr, it, ipa, fpa, C.byref(out), n_threads)
        if rc != 0:
            raise RuntimeError(f"run_op({op}) on {device} failed: " + self.last_error())
        return res

    def model(self, device: str, arch: str, wtype: str = "f16", flags: int = 0, seed: int = 1234, n_threads: int = 0):
        return Model(self, device, arch, wtype, flags, seed, n_threads)


class Model:
    def __init__(self, h: Harness, device, arch, wtype, flags, seed, n_threads):
        self.h = h
        self.lib = h.lib
        self.arch = arch
        self.ptr = self.lib.sdh_model_create(device.encode(), arch.encode(), wtype.encode(), flags, seed, n_threads)
        if not self.ptr:
            raise RuntimeError("sdh_model_create failed: " + h.last_error())

    def close(self):
        if self.ptr:
            self.lib.sdh_model_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def param_bytes(self) -> int:
        return self.lib.sdh_model_param_bytes(self.ptr)

    @property
    def param_count(self) -> int:
        return self.lib.sdh_model_param_count(self.ptr)

    def _args(self, x, t, ctx, y):
        keep = []
        out = []
        for a in (x, t, ctx, y):
            s, arr = _as_sdh(a)
            keep.append(arr)
            out.append(C.byref(s) if s is not None else None)
            keep.append(s)
        return out, keep

    def forward(self, x, t=None, ctx=None, y=None):
        """x: numpy [N,C,H,W]; t: [N]; ctx: [N,77,C]; y: [N,adm].  Returns (out, wall_ms)."""
        (px, pt, pc, py), keep = self._args(x, t, ctx, y)
        ne = (C.c_int64 * 4)()
        self.lib.sdh_model_out_shape(self.ptr, px, ne)
        out = np.empty(tuple(ne)[::-1], np.float32)
        so, _ = _as_sdh(out)
        so.data = out.ctypes.data_as(C.POINTER(C.c_float))
        ms = C.c_double(0)
        rc = self.lib.sdh_model_forward(self.ptr, px, pt, pc, py, C.byref(so), C.byref(ms))
        if rc != 0:
            raise RuntimeError("forward failed: " + self.h.last_error())
        return out, ms.value

    def dump_graph(self, path, x, t=None, ctx=None, y=None):
        (px, pt, pc, py), keep = self._args(x, t, ctx, y)
        n = self.lib.sdh_model_dump_graph(self.ptr, px, pt, pc, py, str(path).encode() if path else None)
        if n < 0:
            raise RuntimeError(self.h.last_error())
        return n, self.lib.sdh_model_last_graph_flops(self.ptr)

    def mailbox_create(self, nbytes: int) -> bytes:
        """Allocate the backend's peer mailbox for payloads of nbytes; returns the 64-byte CUDA IPC handle to hand to the other rank."""
        buf = C.create_string_buffer(64)
        if self.lib.sdh_model_mailbox_create(self.ptr, nbytes, buf) != 0:
            raise RuntimeError(self.h.last_error())
        return buf.raw

    def mailbox_connect(self, peer_handle: bytes | None):
        """Map the other rank's mailbox (None: loopback self test); from now on sample(role=0|1, exchange=None) exchanges on the device."""
        buf = C.create_string_buffer(peer_handle, 64) if peer_handle is not None else None
        if self.lib.sdh_model_mailbox_connect(self.ptr, buf) != 0:
            raise RuntimeError(self.h.last_error())

    def mailbox_close(self):
        self.lib.sdh_model_mailbox_close(self.ptr)

    def add_to_weight(self, name_substr: str, n_dims: int, value: float) -> int:
        """w += value through a graph on the model's backend (the reference's LoRA-apply pattern); returns the element count."""
        n = self.lib.sdh_model_add_to_weight(self.ptr, name_substr.encode(), n_dims, value)
        if n < 0:
            raise RuntimeError(self.h.last_error())
        return n

    def export_graph(self, prefix, x, t=None, ctx=None, y=None) -> int:
        """Write <prefix>.json / <prefix>.bin (graph structure + leaf data) for oracle/graph_f64.py; returns the node count."""
        (px, pt, pc, py), keep = self._args(x, t, ctx, y)
        n = self.lib.sdh_model_export_graph(self.ptr, px, pt, pc, py, str(prefix).encode())
        if n < 0:
            raise RuntimeError(self.h.last_error())
        return n

    def vae_decode(self, z, tile_size: int = 0, overlap: float = 0.5):
        """Reference VAE::decode (optionally with its host-side tiling) -> (image [N,3,8H,8W] in [0,1], wall_ms)."""
        sz, keep = _as_sdh(z)
        ne = (C.c_int64 * 4)()
        self.lib.sdh_model_out_shape(self.ptr, C.byref(sz), ne)
        out = np.empty(tuple(ne)[::-1], np.float32)
        so, _ = _as_sdh(out)
        so.data = out.ctypes.data_as(C.POINTER(C.c_float))
        ms = C.c_double(0)
        self.lib.sdh_vae_decode.restype = C.c_int
        rc = self.lib.sdh_vae_decode(self.ptr, C.byref(sz), C.c_int(tile_size), C.c_float(overlap), C.byref(so), C.byref(ms))
        if rc != 0:
            raise RuntimeError("vae_decode failed: " + self.h.last_error())
        return out, ms.value

    def unsupported_nodes(self, plugin_path, x, t=None, ctx=None, y=None):
        """How many nodes of this model's graph the B200 plugin's supports_op rejects (CPU-only check, no GPU needed)."""
        lib = C.CDLL(str(plugin_path))
        fn = C.cast(lib.ggml_backend_b200_op_supported, C.c_void_p)
        (px, pt, pc, py), keep = self._args(x, t, ctx, y)
        buf = C.create_string_buffer(256)
        self.lib.sdh_model_check_ops.restype = C.c_int
        n = self.lib.sdh_model_check_ops(self.ptr, px, pt, pc, py, fn, buf, C.c_size_t(256))
        if n < 0:
            raise RuntimeError(self.h.last_error())
        return n, buf.value.decode()

    def stats(self) -> dict:
        """Counters of the B200 backend instance behind this model (raises for other backends)."""
        v = (C.c_double * 32)()
        if self.lib.sdh_model_backend_stats(self.ptr, v, 32) != 0:
            raise RuntimeError(self.h.last_error())
        return {k: v[i] for i, k in enumerate(STAT_KEYS)}

    def set_option(self, key: str, value: int):
        if self.lib.sdh_model_set_backend_option(self.ptr, key.encode(), int(value)) != 0:
            raise RuntimeError(f"set_option({key}) failed")

    def sample(self, noise, cond, uncond, steps=20, cfg_scale=7.0, eta=1.0, method="euler_a", sampler_seed=42,
               y_cond=None, y_uncond=None, role=-1, exchange=None):
        sn, a0 = _as_sdh(noise)
        sc, a1 = _as_sdh(cond)
        su, a2 = _as_sdh(uncond)
        syc, a3 = _as_sdh(y_cond)
        syu, a4 = _as_sdh(y_uncond)
        out = np.empty_like(a0)
        so, _ = _as_sdh(out)
        so.data = out.ctypes.data_as(C.POINTER(C.c_float))
        sig = np.zeros(steps + 1, np.float32)
        ts = np.zeros(steps, np.float32)
        nf = C.c_int(0)
        ms = C.c_double(0)
        ref = lambda s: C.byref(s) if s is not None else None
        cb = EXCHANGE_FN(0)
        if exchange is not None:
            def _cb(mine, cond_out, uncond_out, n, user):
                try:
                    a = np.ctypeslib.as_array(mine, shape=(n,))
                    c, u = exchange(a)
                    np.ctypeslib.as_array(cond_out, shape=(n,))[:] = c
                    np.ctypeslib.as_array(uncond_out, shape=(n,))[:] = u
                    return 0
                except Exception as e:  # never let an exception cross the C boundary
                    print("exchange callback failed:", e)
                    return -1
            cb = EXCHANGE_FN(_cb)
        rc = self.lib.sdh_sample_split(self.ptr, method.encode(), steps, cfg_scale, eta, sampler_seed, ref(sn), ref(sc), ref(su),
                                       ref(syc), ref(syu), C.byref(so), sig.ctypes.data_as(C.POINTER(C.c_float)),
                                       ts.ctypes.data_as(C.POINTER(C.c_float)), C.byref(nf), C.byref(ms), role, cb, None)
        if rc != 0:
            raise RuntimeError("sample failed: " + self.h.last_error())
        return out, dict(sigmas=sig, timesteps=ts, n_forwards=nf.value, wall_ms=ms.value)
