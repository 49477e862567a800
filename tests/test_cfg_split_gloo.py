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


WORKER4 = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, r'%(repo)s/stable-diffusion.cpp_b200'); sys.path.insert(0, r'%(repo)s')
    from sdb200 import Harness
    from sdb200.cfg_split import PairExchange
    from oracle.cpu_ref import load_cpu_oracle
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    h = Harness(); load_cpu_oracle(h)
    ex = PairExchange(dist, torch, rank, world, 4 * 16 * 16, 'cpu')     # the class bench.py uses over NCCL
    x = h.randn(42 + 10 * ex.image, (1, 4, 16, 16)); c = h.randn(43, (1, 77, 768)); u = h.randn(44, (1, 77, 768))
    m = h.model('CPU', 'unet_tiny', 'f16', 0, 1234, 1)
    out, info = m.sample(x, c, u, steps=2, cfg_scale=7.0, eta=1.0, role=ex.role, exchange=ex)
    assert info['n_forwards'] == 2 and ex.calls == 2, (info['n_forwards'], ex.calls)
    np.save(os.environ['OUT_PREFIX'] + f'_{rank}.npy', out)
    dist.barrier()
    dist.destroy_process_group()
""")


def test_cfg_split_four_ranks_two_images(tmp_path, cpu_oracle):
    """world_size 4: two independent images, one pair group each (the N=4/8 layout of bench.py)."""
    h = cpu_oracle
    script = tmp_path / "worker4.py"
    script.write_text(WORKER4 % dict(repo=str(REPO)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, OUT_PREFIX=str(tmp_path / "out"), SDH_CPU_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    outs = [np.load(tmp_path / f"out_{r}.npy") for r in range(4)]
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[2], outs[3])
    assert not np.array_equal(outs[0], outs[2]), "the two pairs sample different images"
    c = h.randn(43, (1, 77, 768)); u = h.randn(44, (1, 77, 768))
    m = h.model("CPU", "unet_tiny", "f16", 0, 1234, 1)
    for image in range(2):
        x = h.randn(42 + 10 * image, (1, 4, 16, 16))
        single, _ = m.sample(x, c, u, steps=2, cfg_scale=7.0, eta=1.0)
        assert np.array_equal(outs[2 * image], single), f"image {image}: split sampler differs from the single-process sampler"
    m.close()


def test_batched_cfg_single_forward_per_step(cpu_oracle):
    """role 2: cond + uncond as ONE N = 2 forward per step (the other way to use the CFG batch, on a single device).  Same host
    sampler; the reference CPU backend's GEMM blocking depends on the batch, so the latent agrees to rounding noise amplified by the
    ancestral sampler, not bit for bit."""
    h = cpu_oracle
    x = h.randn(42, (1, 4, 16, 16)); c = h.randn(43, (1, 77, 768)); u = h.randn(44, (1, 77, 768))
    m = h.model("CPU", "unet_tiny", "f16", 0, 1234, 2)
    serial, i0 = m.sample(x, c, u, steps=2, cfg_scale=7.0, eta=1.0)
    batched, i1 = m.sample(x, c, u, steps=2, cfg_scale=7.0, eta=1.0, role=2)
    m.close()
    assert i0["n_forwards"] == 4 and i1["n_forwards"] == 2
    assert np.array_equal(i0["sigmas"], i1["sigmas"]) and np.array_equal(i0["timesteps"], i1["timesteps"])
    r = np.linalg.norm(serial - batched) / np.linalg.norm(serial)
    assert r < 1e-2, f"rel_l2 {r:.2e}"
