"""CPU-oracle vs B200 whole-model comparison through the harness (run on a GPU box)."""
import json, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "stable-diffusion.cpp_b200")); sys.path.insert(0, str(REPO))
from sdb200 import Harness
from oracle.cpu_ref import load_cpu_oracle

def rel_l2(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b.astype(np.float64)) + 1e-30))

def main():
    cases = sys.argv[1:] or ["unet_tiny:f16:0:32", "unet_tiny:f16:1:32", "vae_decoder:f16:0:16", "sd15_unet:f16:0:64", "sd15_unet:f16:1:64"]
    h = Harness(); var = load_cpu_oracle(h); devs = h.load_b200()
    print("devices", h.devices(), "cpu variant", var)
    results = []
    for case in cases:
        arch, wtype, flags, size = case.split(":"); flags = int(flags); size = int(size)
        x = h.randn(42, (1, 4, size, size)); ctx = h.randn(43, (1, 77, 768)); t = np.array([999.0], np.float32)
        args = (x,) if arch.startswith("vae") else (x, t, ctx)
        outs = {}
        for dev in ("CPU", devs[0]):
            m = h.model(dev, arch, wtype, flags, 1234, 0)
            out, ms = m.forward(*args)
            t0 = time.time(); out2, ms2 = m.forward(*args); wall = (time.time() - t0) * 1e3
            outs[dev] = out2
            print(f"  {case} {dev}: first {ms:.1f} ms, second {ms2:.1f} ms, rms {out2.std():.4f} finite {np.isfinite(out2).all()} repeat-identical {np.array_equal(out, out2)}")
            m.close()
        r = rel_l2(outs[devs[0]], outs["CPU"])
        d = np.abs(outs[devs[0]] - outs["CPU"]).max()
        print(f"{case}: rel_l2 {r:.3e} max_abs {d:.3e}")
        results.append(dict(case=case, rel_l2=r, max_abs=float(d)))
    (REPO / "gpurun_out" / "model_parity.json").write_text(json.dumps(results, indent=1))

if __name__ == "__main__":
    main()
