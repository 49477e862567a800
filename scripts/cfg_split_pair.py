"""Two processes, one per GPU (torchrun --nproc-per-node 2): CFG split with the device-side exchange over NVLink peer memory.
Rank r evaluates branch r (0 = cond, 1 = uncond) of the same image; rank 0 also samples the image serially on its own GPU and both
latents must be BIT-IDENTICAL.  Prints one JSON line on rank 0.   usage: cfg_split_pair.py [arch] [steps]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "stable-diffusion.cpp_b200"))
sys.path.insert(0, str(REPO))


def main():
    import torch
    import torch.distributed as dist
    arch = sys.argv[1] if len(sys.argv) > 1 else "unet_tiny"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")              # host-side plumbing only: exchanges the two 64-byte IPC handles
    from sdb200 import Harness, FLAG_FLASH_ATTN
    h = Harness()
    h.load_b200()
    dev = f"B200_{local}"
    shape = (1, 4, 16, 16) if arch == "unet_tiny" else (1, 4, 64, 64)
    x = h.randn(42, shape); c = h.randn(43, (1, 77, 768)); u = h.randn(44, (1, 77, 768))
    m = h.model(dev, arch, "f16", FLAG_FLASH_ATTN, 1234, 0)
    serial = None
    if rank == 0:
        serial, _ = m.sample(x, c, u, steps=steps, cfg_scale=7.0, eta=1.0)
    mine = torch.frombuffer(bytearray(m.mailbox_create(int(np.prod(shape)) * 4)), dtype=torch.uint8).clone()
    handles = [torch.zeros(64, dtype=torch.uint8) for _ in range(2)]
    dist.all_gather(handles, mine)
    m.mailbox_connect(bytes(handles[1 - rank].numpy().tobytes()))
    dist.barrier()
    m.sample(x, c, u, steps=2, cfg_scale=7.0, eta=1.0, role=rank)            # warm-up: eager + capture
    dist.barrier()
    t0 = time.perf_counter()
    s0 = m.stats()
    split, info = m.sample(x, c, u, steps=steps, cfg_scale=7.0, eta=1.0, role=rank)
    wall = time.perf_counter() - t0
    s1 = m.stats()
    both = [torch.zeros(split.size, dtype=torch.float32) for _ in range(2)]
    dist.all_gather(both, torch.from_numpy(split.reshape(-1).copy()))
    if rank == 0:
        same_ranks = bool(torch.equal(both[0], both[1]))
        print(json.dumps(dict(arch=arch, steps=steps, bit_identical_to_serial=bool(np.array_equal(split, serial)), ranks_agree=same_ranks,
                              forwards_per_rank=info["n_forwards"], peer_exchanges=int(s1["peer_exchanges"] - s0["peer_exchanges"]),
                              fused_epilogue_pushes=int(s1["cta2_gemm_launches"] - s0["cta2_gemm_launches"]),
                              cuda_graph_replays=int(s1["cuda_graph_replays"] - s0["cuda_graph_replays"]),
                              device_ms_per_step=(s1["total_graph_ms"] - s0["total_graph_ms"]) / steps, wall_ms_per_step=1e3 * wall / steps)), flush=True)
    m.mailbox_close()
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
