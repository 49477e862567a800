"""Run N forwards of one harness architecture on B200_0 (for `ncu` launch lists / full captures of the non-headline configs).
usage: one_forward.py <sdxl|flux|sd15|sd15x2|mmdit|wan|vae|vae128> [n_forwards]        env GGML_B200_CUDA_GRAPHS=0 keeps every launch visible"""
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "stable-diffusion.cpp_b200"))
from sdb200 import Harness  # noqa: E402

CASES = {
    "sd15": ("sd15_unet", "f16", dict(x=(1, 4, 64, 64), ctx=(1, 77, 768), y=None, t=999.0)),
    "sd15x2": ("sd15_unet", "f16", dict(x=(2, 4, 64, 64), ctx=(2, 77, 768), y=None, t=999.0)),
    "sdxl": ("sdxl_unet", "bf16", dict(x=(1, 4, 128, 128), ctx=(1, 77, 2048), y=(1, 2816), t=999.0)),
    "flux1": ("flux_1x1", "bf16", dict(x=(1, 16, 128, 128), ctx=(1, 256, 4096), y=(1, 768), t=1.0)),
    "flux": ("flux_schnell", "bf16", dict(x=(1, 16, 128, 128), ctx=(1, 256, 4096), y=(1, 768), t=1.0)),
    "mmdit": ("mmdit_sd3", "f16", dict(x=(1, 16, 128, 128), ctx=(1, 154, 4096), y=(1, 2048), t=500.0)),
    "vae": ("vae_decoder", "f16", dict(x=(1, 4, 64, 64), ctx=None, y=None, t=None)),
    "vae128": ("vae_decoder", "f16", dict(x=(1, 4, 128, 128), ctx=None, y=None, t=None)),
    "wan": ("wan_1_3b", "q8_0", dict(x=(16, 13, 64, 64), ctx=(1, 512, 4096), y=None, t=500.0)),
}


def main():
    name = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    arch, wtype, e = CASES[name]
    h = Harness()
    dev = h.load_b200()[0]
    m = h.model(dev, arch, wtype, 1, 1234, 0)
    x = h.randn(42, e["x"]); ctx = h.randn(43, e["ctx"]) if e["ctx"] else None; y = h.randn(44, e["y"]) if e["y"] else None
    t = np.full((e["x"][0] if name == "sd15x2" else 1,), e["t"], np.float32) if e["t"] is not None else None
    nodes, flops = m.dump_graph(None, x, t, ctx, y)
    for i in range(n):
        s0 = m.stats()
        t0 = time.perf_counter()
        out, _ = m.forward(x, t, ctx, y)
        wall = (time.perf_counter() - t0) * 1e3
        s1 = m.stats()
        d = s1["total_graph_ms"] - s0["total_graph_ms"]
        print(f"{name} forward {i}: device {d:.2f} ms, wall {wall:.1f} ms, {flops / 1e12:.2f} TFLOP -> {flops / 1e9 / max(d, 1e-9):.1f} TFLOP/s, "
              f"{int(s1['kernel_launches'] - s0['kernel_launches'])} launches, {nodes} nodes, finite {bool(np.isfinite(out).all())}", flush=True)
    m.close()


if __name__ == "__main__":
    main()
