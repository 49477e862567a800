"""oracle/cpu_ref.py -- TEST INFRASTRUCTURE (checker only; never on the product path).

Registers the reference's own ggml CPU backend (compiled from /root/reference by
oracle/Makefile into oracle/_ref/) in a harness registry.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from pathlib import Path

ORACLE_DIR = Path(__file__).resolve().parent / "_ref"
# best-first; ggml's loader refuses a variant whose ggml_backend_score() is 0 on this host CPU
# (ggml-backend-reg.cpp:221-266, ggml-cpu/arch/x86/cpu-feats.cpp)
VARIANTS = ("sapphirerapids", "skylakex", "haswell")


def load_cpu_oracle(harness) -> str:
    """Load the best CPU variant the host supports; returns the variant name."""
    if "CPU" in harness.devices():
        return "already-loaded"
    for v in VARIANTS:
        so = ORACLE_DIR / f"libggml-cpu-{v}.so"
        if so.exists() and harness.load_backend(so) > 0 and "CPU" in harness.devices():
            return v
    raise RuntimeError(f"no usable CPU oracle variant under {ORACLE_DIR} (run `make -C oracle` where /root/reference exists)")


def test_backend_ops() -> Path:
    return ORACLE_DIR / "test-backend-ops"


def usable_cores() -> int:
    """Host cores this process may actually use: min(affinity mask, cgroup cpu.max quota).  os.cpu_count() alone
    over-reports inside containers (a 128-thread ggml pool on an 8-core quota runs ~30x slower than 8 threads)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except Exception:
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = min(n, max(1, int(quota / period + 0.5)))
        except Exception:
            pass
    return max(1, n)


def best_thread_count(harness, budget_s: float = 20.0) -> int:
    """Pick the ggml CPU thread count that is fastest on a tiny UNet forward ("all the host threads it can use" without
    oversubscribing): tries usable_cores() and its halvings, stops as soon as more threads stop helping."""
    import time
    import numpy as np
    x = harness.randn(1, (1, 4, 16, 16)); ctx = harness.randn(2, (1, 77, 768)); t = np.array([500.0], np.float32)
    best, best_t = 1, float("inf")
    n = usable_cores()
    t_start = time.time()
    cands = []
    c = min(4, n)
    while c < n:
        cands.append(c)
        c *= 2
    cands.append(n)          # ascending: an oversubscribed pool is ~30x slower, so never START with the largest count
    for c in cands:
        if time.time() - t_start > budget_s:
            break
        m = harness.model("CPU", "unet_tiny", "f16", 0, 1, c)
        m.forward(x, t, ctx)
        t0 = time.time(); m.forward(x, t, ctx); dt = time.time() - t0
        m.close()
        if dt < best_t:
            best, best_t = c, dt
        elif dt > 1.5 * best_t:
            break
    return best
