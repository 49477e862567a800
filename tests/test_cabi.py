"""CPU-only: the drop-in boundary loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import re
from pathlib import Path

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
