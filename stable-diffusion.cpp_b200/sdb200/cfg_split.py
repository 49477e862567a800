"""CFG-batch split across GPU pairs (SURVEY.md 8e level 1).

Ranks (2i, 2i+1) evaluate the cond / uncond forward of image i and exchange the eps prediction with ONE all-gather per
step inside their pair; the guidance mix and the sampler update are then done redundantly by both ranks with the
reference's own host code (`sdh_sample_split`).  The reference evaluates the two forwards serially on one device
(`src/stable-diffusion.cpp:2811-2836`); this is the minimal caller-side change that overlaps them.

`torch.distributed` is plumbing only.  The same class runs on `cuda` tensors over NCCL (bench.py) and on CPU tensors over
gloo (tests/test_cfg_split_gloo.py).
"""
from __future__ import annotations

import numpy as np


class PairExchange:
    """`exchange(mine) -> (cond_eps, uncond_eps)` callback for `Model.sample(role=rank % 2, exchange=...)`."""

    def __init__(self, dist, torch, rank: int, world: int, n_elems: int, device: str = "cuda"):
        if world < 2 or world % 2:
            raise ValueError(f"CFG split needs an even world size >= 2, got {world}")
        self.dist, self.torch = dist, torch
        self.image = rank // 2
        self.role = rank % 2
        self.cuda = device.startswith("cuda")
        self.coll_ms = 0.0
        self.calls = 0
        # new_group is collective over the whole world: every rank creates every pair group, in the same order
        if world > 2:
            groups = [dist.new_group(ranks=[2 * i, 2 * i + 1]) for i in range(world // 2)]
            self.group = groups[self.image]
        else:
            self.group = None
        self.send = torch.empty(n_elems, dtype=torch.float32, device=device)
        self.recv = torch.empty(2 * n_elems, dtype=torch.float32, device=device)
        if self.cuda:
            self.host = torch.empty(2 * n_elems, dtype=torch.float32).pin_memory()
            self.ev0 = torch.cuda.Event(enable_timing=True)
            self.ev1 = torch.cuda.Event(enable_timing=True)

    def __call__(self, mine: np.ndarray):
        torch, dist = self.torch, self.dist
        n = mine.size
        self.calls += 1
        self.send.copy_(torch.from_numpy(np.ascontiguousarray(mine).reshape(-1)), non_blocking=True)
        if self.cuda:
            self.ev0.record()
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)   # THE collective of the path: 64 KB over NVLink
            self.ev1.record()
            self.host.copy_(self.recv, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self.coll_ms += self.ev0.elapsed_time(self.ev1)
            both = self.host.numpy()
        else:
            parts = [self.recv[:n], self.recv[n:]]
            dist.all_gather(parts, self.send, group=self.group)
            both = self.recv.numpy()
        return both[:n], both[n:]
