"""Run unet_tiny three times on B200 under different backend options; report run-to-run equality and error vs CPU."""
import os, sys, subprocess, json
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, r"%s/stable-diffusion.cpp_b200"); sys.path.insert(0, r"%s")
from sdb200 import Harness
from oracle.cpu_ref import load_cpu_oracle
h = Harness(); load_cpu_oracle(h); dev = h.load_b200()[0]
arch, fa = sys.argv[1], int(sys.argv[2])
x = h.randn(42, (1, 4, 16, 16)); ctx = h.randn(43, (1, 77, 768)); t = np.array([999.0], np.float32)
m = h.model("CPU", arch, "f16", fa, 1234, 0); ref, _ = m.forward(x, t, ctx); m.close()
m = h.model(dev, arch, "f16", fa, 1234, 0)
outs = [m.forward(x, t, ctx)[0] for _ in range(4)]
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
print("RESULT", [round(rel(o, ref), 5) for o in outs], [bool(np.array_equal(outs[0], o)) for o in outs], m.stats()["cuda_graph_replays"])
''' % (REPO, REPO)
combos = [dict(), dict(GGML_B200_CUDA_GRAPHS="0"), dict(GGML_B200_FUSION="0"), dict(GGML_B200_CUDA_GRAPHS="0", GGML_B200_FUSION="0"),
          dict(GGML_B200_IMPLICIT_CONV="0"), dict(GGML_B200_FUSED_ATTN="0")]
for fa in (0, 1):
    for c in combos:
        env = dict(os.environ, **c)
        r = subprocess.run([sys.executable, "-c", CHILD, "unet_tiny", str(fa)], env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(f"fa={fa} {c}: {line[0] if line else 'FAILED ' + r.stderr[-300:]}")
