"""GPU: the CFG-batch split's device-side exchange (kernels/peer.cu, include/ggml-b200.h ggml_backend_b200_peer_mailbox_*).

  loopback   ONE GPU: the rank is its own peer.  Exercises the whole mechanism -- mailbox, the convolution epilogue that stores the
             eps prediction into the (here: own) mailbox, sequence flags, the device-side wait, CUDA-graph capture and replay of all
             of it -- and must reproduce the serial sampler bit for bit when cond == uncond.
  pair       TWO GPUs, two processes (torchrun): rank r evaluates branch r; the latent must be bit-identical to the serial sampler
             on one GPU.  Skipped when the box has a single GPU."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]


def test_loopback_exchange_reproduces_the_serial_sampler(b200):
    h, dev = b200
    x = h.randn(42, (1, 4, 16, 16)); c = h.randn(43, (1, 77, 768))
    m = h.model(dev, "unet_tiny", "f16", 1, 1234, 0)
    serial, i0 = m.sample(x, c, c, steps=4, cfg_scale=7.0, eta=1.0)          # cond == uncond: both branches compute the same eps
    m.mailbox_create(x.size * 4)
    m.mailbox_connect(None)
    s0 = m.stats()
    split, i1 = m.sample(x, c, c, steps=4, cfg_scale=7.0, eta=1.0, role=0)
    s1 = m.stats()
    m.mailbox_close()
    again, _ = m.sample(x, c, c, steps=4, cfg_scale=7.0, eta=1.0)
    m.close()
    assert i0["n_forwards"] == 8 and i1["n_forwards"] == 4
    assert s1["peer_exchanges"] - s0["peer_exchanges"] == 4
    assert s1["cuda_graph_replays"] - s0["cuda_graph_replays"] >= 2, "the exchange must live inside the replayed CUDA graph"
    print("output convolution on the CTA-pair kernel (fused peer store):", s1["cta2_gemm_launches"] - s0["cta2_gemm_launches"] >= 4)
    assert np.array_equal(split, serial)
    assert np.array_equal(again, serial)


def _gpu_count():
    try:
        import ctypes
        from sdb200 import B200_SO
        lib = ctypes.CDLL(str(B200_SO))
        lib.ggml_backend_b200_get_device_count.restype = ctypes.c_int
        return lib.ggml_backend_b200_get_device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs on the box")
@pytest.mark.parametrize("arch", ["unet_tiny", "sd15_unet"])
def test_pair_exchange_is_bit_identical_to_serial(arch):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29533", str(REPO / "scripts" / "cfg_split_pair.py"), arch, "4"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-3000:]
    d = json.loads(lines[-1])
    assert d["bit_identical_to_serial"] and d["ranks_agree"], d
    assert d["peer_exchanges"] == 4 and d["forwards_per_rank"] == 4
