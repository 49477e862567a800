"""CPU-only: the drop-in boundary loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]


def _declared_functions(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+;", "", text, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{}]*\)\s*;", text)
    return sorted({n for n in names if n.startswith(("ggml_backend_", "sdh_"))})


def test_plugin_exports_declared_cabi():
    from sdb200 import B200_SO
    assert B200_SO.exists(), "libggml-b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(str(B200_SO))
    names = _declared_functions(REPO / "include" / "ggml-b200.h")
    assert "ggml_backend_init" in names and "ggml_backend_score" in names and len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"symbol {n} declared in include/ggml-b200.h is not exported"


def test_plugin_registry_shape_without_compute():
    """No compute calls: registry object, api version, device count consistent with score."""
    from sdb200 import B200_SO
    lib = ctypes.CDLL(str(B200_SO))
    lib.ggml_backend_init.restype = ctypes.c_void_p
    lib.ggml_backend_b200_reg.restype = ctypes.c_void_p
    lib.ggml_backend_score.restype = ctypes.c_int
    lib.ggml_backend_b200_get_device_count.restype = ctypes.c_int
    reg = lib.ggml_backend_init()
    assert reg and reg == lib.ggml_backend_b200_reg()
    api_version = ctypes.c_int.from_address(reg).value          # struct ggml_backend_reg { int api_version; ... }
    assert api_version == 2                                      # GGML_BACKEND_API_VERSION, ggml-backend-impl.h:11
    n = lib.ggml_backend_b200_get_device_count()
    assert (lib.ggml_backend_score() > 0) == (n > 0)
    lib.ggml_backend_b200_init.restype = ctypes.c_void_p
    assert lib.ggml_backend_b200_init(10_000) is None            # invalid device -> NULL, not a crash


def test_harness_exports_declared_cabi():
    from sdb200 import HARNESS_SO
    assert HARNESS_SO.exists()
    lib = ctypes.CDLL(str(HARNESS_SO))
    names = _declared_functions(REPO / "include" / "sd_b200_harness.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"symbol {n} declared in include/sd_b200_harness.h is not exported"


def test_loader_accepts_or_skips_plugin(harness):
    """The reference's own loader (ggml_backend_load, ggml-backend-reg.cpp:221-266) must either register our devices
    (GPU box) or refuse politely because ggml_backend_score() == 0 (CPU box) -- never crash."""
    from sdb200 import B200_SO
    before = harness.devices()
    harness.load_backend(B200_SO)
    after = harness.devices()
    assert set(before) <= set(after)
    lib = ctypes.CDLL(str(B200_SO))
    lib.ggml_backend_b200_get_device_count.restype = ctypes.c_int
    assert len([d for d in after if d.startswith("B200_")]) == lib.ggml_backend_b200_get_device_count()


def test_product_path_has_no_cpu_fallback(harness):
    """Asking for a model on a device that does not exist must fail loudly."""
    with pytest.raises(RuntimeError):
        harness.model("B200_99", "unet_tiny", "f16", 0, 1, 1)


@pytest.mark.parametrize("arch,wtype,flags,xs,cs,ys", [
    ("unet_tiny", "f16", 0, (1, 4, 16, 16), (1, 77, 768), None),
    ("unet_tiny", "f16", 1, (2, 4, 16, 16), (2, 77, 768), None),            # flash-attention graph, batched CFG (N = 2)
    ("sd15_unet", "f16", 1, (1, 4, 64, 64), (1, 77, 768), None),
    ("vae_decoder", "f16", 0, (1, 4, 16, 16), None, None),
    ("flux_tiny", "bf16", 1, (1, 16, 32, 32), (1, 64, 4096), (1, 768)),
    ("wan_1_3b", "q8_0", 1, (16, 3, 16, 16), (1, 512, 4096), None),
    ("sdxl_unet", "bf16", 1, (1, 4, 32, 32), (1, 77, 2048), (1, 2816)),
    ("mmdit_sd3", "f16", 1, (1, 16, 32, 32), (1, 154, 4096), (1, 2048)),
    ("wan_vae_decoder", "f16", 0, (16, 2, 8, 8), None, None),                # causal 3-D conv decoder: IM2COL_3D, PAD, RMS_NORM, CONCAT caches
])
def test_every_graph_node_is_claimed_by_the_plugin(harness, arch, wtype, flags, xs, cs, ys):
    """No silent CPU fallback, checked WITHOUT a GPU: build the reference's graph for each model family (on the CPU device, nothing is
    computed) and ask the plugin's supports_op about every node.  sd.cpp switches a runner to ggml_backend_sched with a CPU fallback
    backend as soon as one node is refused (src/core/ggml_extend.hpp:2198-2225); the north star forbids that."""
    from sdb200 import B200_SO
    from oracle.cpu_ref import load_cpu_oracle
    load_cpu_oracle(harness)
    m = harness.model("CPU", arch, wtype, flags | 4, 1234, 2)          # bit 2: parameters placed, not filled -- nothing is computed here
    x = harness.randn(42, xs)
    t = np.full((xs[0] if arch == "unet_tiny" else 1,), 999.0, np.float32) if cs is not None or arch != "vae_decoder" else None
    ctx = harness.randn(43, cs) if cs else None
    y = harness.randn(44, ys) if ys else None
    bad, first = m.unsupported_nodes(B200_SO, x, t, ctx, y)
    m.close()
    assert bad == 0, f"{arch}: {bad} node(s) would fall back to the CPU, first: {first}"


def test_clip_text_encoder_graph_is_claimed_by_the_plugin(harness):
    """SURVEY.md 8f-2 (the stage right before the hot path): the CLIP ViT-L/14 text encoder graph of the reference (GET_ROWS embedding lookup,
    LayerNorm, causal-mask attention, quick-GELU MLP) consists only of ops the plugin executes -- a GPU-resident txt2img needs no CPU
    fallback for it either."""
    from sdb200 import B200_SO
    from oracle.cpu_ref import load_cpu_oracle
    load_cpu_oracle(harness)
    m = harness.model("CPU", "clip_l", "f16", 0, 1234, 2)
    ids = np.full((1, 1, 1, 77), 49407, np.float32)
    ids[0, 0, 0, 0] = 49406
    ids[0, 0, 0, 1:9] = [320, 1125, 539, 2368, 525, 1929, 267, 1662]
    out, _ = m.forward(ids)
    bad, first = m.unsupported_nodes(B200_SO, ids)
    m.close()
    assert out.shape[-2:] == (77, 768) and np.isfinite(out).all()
    assert bad == 0, f"{bad} node(s) would fall back to the CPU, first: {first}"


def test_t5_text_encoder_graph_is_claimed_by_the_plugin(harness):
    """the T5-XXL encoder of Flux / SD3 (src/model/te/t5.hpp) at full width (4 of its 24 layers): RMS-style layer norm,
    relative-position-bias attention (GET_ROWS of the bucket table), gated-GELU feed forward -- every node is claimed by the plugin."""
    from sdb200 import B200_SO
    from oracle.cpu_ref import load_cpu_oracle
    load_cpu_oracle(harness)
    m = harness.model("CPU", "t5_xxl_4l", "f16", 0, 1234, 4)
    ids = np.zeros((1, 1, 1, 32), np.float32)
    ids[0, 0, 0, :8] = [71, 1712, 13, 3, 9, 1782, 5, 1]
    out, _ = m.forward(ids)
    bad, first = m.unsupported_nodes(B200_SO, ids)
    m.close()
    assert out.shape[-2:] == (32, 4096) and np.isfinite(out).all()
    assert bad == 0, f"{bad} node(s) would fall back to the CPU, first: {first}"


def test_halo_conv_plan_model_against_committed_hardware_sweep():
    """The fitted cost model of the halo-reuse convolution (gemm_tc.cu conv_halo_model) must keep choosing plans within 15 % of the best
    plan MEASURED on B200 for every shape of the committed sweep (profiles/r02_conv_halo_sweep.log: tile width x split-K x taps per box,
    each line element-exact against the per-tap kernel), and its predicted time within 35 % of the measured one."""
    from sdb200 import B200_SO
    lib = ctypes.CDLL(str(B200_SO))
    fn = lib.ggml_backend_b200_debug_conv_plan
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int64] * 5 + [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3 + [ctypes.POINTER(ctypes.c_double)]
    shapes = {"sd15 64x64 320->320 x2": (2, 64, 64, 320, 320), "sd15 32x32 640->640 x2": (2, 32, 32, 640, 640),
              "sd15 16x16 1280->1280 x2": (2, 16, 16, 1280, 1280), "sd15 64x64 640->320 x2": (2, 64, 64, 640, 320),
              "sd15 32x32 1280->640 x2": (2, 32, 32, 1280, 640), "vae 512x512 128->128": (1, 512, 512, 128, 128),
              "vae 256x256 256->256": (1, 256, 256, 256, 256), "vae 128x128 512->512": (1, 128, 128, 512, 512),
              "sdxl 128x128 320->320": (1, 128, 128, 320, 320), "sdxl 64x64 640->640": (1, 64, 64, 640, 640)}
    meas = {}
    for line in (REPO / "profiles" / "r02_conv_halo_sweep.log").read_text().splitlines():
        m = re.match(r"(\S+ \S+ \S+(?: x2)?)\s+(\d+)\s+(\d+)\s+(\d+) \|\s+([\d.]+)", line)
        if m:
            meas[(m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)))] = float(m.group(5))
    assert len(meas) > 150
    for name, (n, H, W, C, OC) in shapes.items():
        bn, sp, taps, us = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        assert fn(n, H, W, C, OC, 148, ctypes.byref(bn), ctypes.byref(sp), ctypes.byref(taps), ctypes.byref(us)) == 1, name
        key = (name, bn.value, sp.value, taps.value)
        assert key in meas, f"{name}: chosen plan {key[1:]} was not part of the hardware sweep"
        best = min(v for k, v in meas.items() if k[0] == name and k[3] > 0)
        assert meas[key] <= 1.15 * best, f"{name}: chosen {key[1:]} measured {meas[key]} us, best halo plan {best} us"
        assert abs(us.value - meas[key]) <= 0.35 * meas[key], f"{name}: model {us.value:.1f} us vs measured {meas[key]} us"
    assert fn(1, 8, 8, 1280, 1280, 148, None, None, None, None) == 0      # 8 x 8 level: outside the 16 x 8 patch envelope


def test_pair_kernel_geometry_invariants_over_every_legal_plan():
    """fill_geometry (gemm_tc2.cu) for every tile width x split-K x taps-per-box the dispatchers can ask for: the invariants the kernel's
    addressing relies on -- stages of whole 1024-byte swizzle atoms, the ring + the two 16 KB staging tiles inside the 227 KB of dynamic shared
    memory, a split-K partial tile [bn][128] f32 that fits the ring it reuses, two accumulators of bn columns inside a power-of-two TMEM
    allocation of at most 512 columns, and a grid of whole CTA pairs that never exceeds the SM count in the persistent (split 1) form."""
    from sdb200 import B200_SO
    lib = ctypes.CDLL(str(B200_SO))
    fn = lib.ggml_backend_b200_debug_pair_geometry
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int)] * 2 + [ctypes.POINTER(ctypes.c_int64)] + [ctypes.POINTER(ctypes.c_int)] * 2
    legal = 0
    for taps in (0, 3, 9):
        for bn in range(16, 257, 16):
            for splits in (1, 2, 4):
                for (M, N, batch, nkb) in ((262144, 128, 1, 18), (4096, 320, 2, 45), (256, 1280, 2, 180), (1024, 640, 2, 90), (320, 4096, 2, 5)):
                    st, sb, tm, ct = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
                    sm = ctypes.c_int64()
                    if splits > nkb:
                        continue
                    if not fn(M, N, batch, nkb, bn, splits, taps, 148, ctypes.byref(st), ctypes.byref(sb), ctypes.byref(sm), ctypes.byref(tm), ctypes.byref(ct)):
                        continue
                    legal += 1
                    tag = (taps, bn, splits, M, N)
                    assert sb.value % 1024 == 0, tag
                    a_bytes = {0: 16384, 3: 20 * 1024, 9: 23 * 1024}[taps]
                    assert sb.value == a_bytes + max(taps, 1) * (bn // 2) * 128, tag
                    assert 2 <= st.value <= 10 and (taps != 3 or st.value >= 3), tag
                    assert sm.value <= 227 * 1024 - 2048, tag
                    assert sm.value >= st.value * sb.value + (32768 if splits == 1 else 0), tag
                    if splits > 1:
                        assert st.value * sb.value >= bn * 128 * 4, tag
                    assert tm.value in (32, 64, 128, 256, 512) and tm.value >= 2 * bn, tag
                    assert ct.value % 2 == 0 and ct.value >= 2, tag
                    if splits == 1:
                        assert ct.value <= 148, tag
    assert legal > 300
    assert fn(4096, 320, 2, 45, 24, 1, 0, 148, None, None, None, None, None) == 0          # tile width must be a multiple of 16
    assert fn(4096, 320, 2, 45, 256, 1, 9, 148, None, None, None, None, None) == 0         # nine 16 KB filter tiles + the image box: no two stages fit
