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
