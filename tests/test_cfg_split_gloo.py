"""CPU, world_size 2, gloo: the N>1 path of bench.py -- CFG batch split with one all-gather of the eps prediction per
step -- must reproduce the single-process sampler exactly (same host code, same weights; only the placement differs)."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]

WORKER = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, r'%(repo)s/stable-diffusion.cpp_b200'); sys.path.insert(0, r'%(repo)s')
    from sdb200 import Harness
    from oracle.cpu_ref import load_cpu_oracle
    dist.init_process_group('gloo')
    rank = dist.get_rank()
    h = Harness(); load_cpu_oracle(h)
    x = h.randn(42, (1, 4, 16, 16)); c = h.randn(43, (1, 77, 768)); u = h.randn(44, (1, 77, 768))
    m = h.model('CPU', 'unet_tiny', 'f16', 0, 1234, 2)
    calls = [0]
    def exchange(mine):
        calls[0] += 1
        parts = [torch.empty(mine.size, dtype=torch.float32) for _ in range(2)]
        dist.all_gather(parts, torch.from_numpy(mine.copy()))
        return parts[0].numpy(), parts[1].numpy()
    out, info = m.sample(x, c, u, steps=3, cfg_scale=7.0, eta=1.0, role=rank, exchange=exchange)
    assert info['n_forwards'] == 3 and calls[0] == 3, (info['n_forwards'], calls[0])
    np.save(os.environ['OUT_PREFIX'] + f'_{rank}.npy', out)
    dist.destroy_process_group()
""")


def test_cfg_split_two_ranks_matches_single_process(tmp_path, cpu_oracle):
    h = cpu_oracle
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(repo=str(REPO)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, OUT_PREFIX=str(tmp_path / "out"))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    o0, o1 = np.load(tmp_path / "out_0.npy"), np.load(tmp_path / "out_1.npy")
    assert np.array_equal(o0, o1), "both ranks of a pair must hold the same latent"
    gold = np.load(REPO / "tests" / "golden" / "cpu_models.npz")["unet_tiny_sample3"]
    x = h.randn(42, (1, 4, 16, 16)); c = h.randn(43, (1, 77, 768)); u = h.randn(44, (1, 77, 768))
    m = h.model("CPU", "unet_tiny", "f16", 0, 1234, 2)
    single, _ = m.sample(x, c, u, steps=3, cfg_scale=7.0, eta=1.0)
    m.close()
    assert np.array_equal(o0, single), "split sampler differs from the single-process sampler"
    assert np.linalg.norm(o0 - gold) / np.linalg.norm(gold) < 1e-4
