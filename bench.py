#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, through the reference-facing boundary.

Workload (config.workload = "sd15_txt2img_512_euler_a_cfg7", BASELINE.json configs[1]): SD1.5 UNet, latent 64x64x4
(512x512 image), context 77x768, F16 weights (synthetic, seeded: no checkpoints exist offline), flash-attention graph
variant, Euler-a sampling with CFG 7.0.  One "step" = one denoise step of one image = the TWO UNet forwards (cond +
uncond) that the reference's sample() issues per sigma plus its host-side sampler math -- all of it the reference's own
unmodified host code (host/_ref/libsd_harness.so), dispatching through ggml's backend vtable into libggml-b200.so.

  value  denoise steps/s from DEVICE time: CUDA events on the backend stream around every graph_compute of the K timed
         steps (inputs are resident in HBM at that point), max over ranks
  e2e    the same K steps timed by wall clock around the reference sampler call: graph build, gallocr, H2D of x / t /
         context, graph_compute, D2H of the eps prediction, CFG combine and Euler-a update on the host
  N > 1  CFG batch split (north_star / SURVEY.md 8e): ranks (2i, 2i+1) evaluate cond / uncond of image i and all-gather the
         eps prediction over NCCL (64 KB) every step; N/2 images run concurrently (N = 1: one GPU does both branches)
  --impl reference   the reference's own CPU implementation (ggml CPU backend compiled from /root/reference into
         oracle/_ref) on the same workload with all host threads, bounded to a few minutes
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO / "stable-diffusion.cpp_b200"))
sys.path.insert(0, str(REPO))

WORKLOAD = "sd15_txt2img_512_euler_a_cfg7"
FLOPS_PER_FORWARD = 0.80327e12        # algorithmic, recomputed from the live graph below
CFG_SCALE, ETA = 7.0, 1.0


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"], src="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


LAYOUT_TEXT = {
    "batched": "every GPU samples its own image, cond + uncond as ONE N = 2 UNet forward per step (caller-side CFG batching; no collective)",
    "serial": "the reference's own order: two N = 1 UNet forwards per step on one GPU (stable-diffusion.cpp:2811-2836)",
    "cfg-split": "CFG batch split over GPU pairs: rank 2i evaluates cond, rank 2i+1 uncond of image i; the eps prediction (64 KB) is exchanged once per "
                 "step by device code over NVLink peer memory (stored into the partner's HBM by the output convolution's epilogue, flag handshake, "
                 "inside the replayed CUDA graph: kernels/peer.cu) -- the north star's single all-gather on the latent, without NCCL or host staging",
}


def ncu_traffic():
    """DRAM bytes (read + write) per tcgen05 GEMM launch: the sum over EVERY GEMM launch of one batched-CFG forward divided by their
    count, from the committed whole-forward ncu metrics pass (profiles/r02_ncu_metrics_sd15_batched.json, produced by
    scripts/ncu_kernel_table.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,...` over all launches of the forward).
    Offline evidence, never measured under the timed run; null when absent."""
    for name in ("r02_ncu_metrics_sd15_batched.json",):
        p = REPO / "profiles" / name
        if not p.exists():
            continue
        d = json.loads(p.read_text())
        rows = [k for k in d.get("kernels", []) if "k_gemm_tc" in k["kernel"]]
        n = sum(k["launches"] for k in rows)
        if n:
            tot = sum(k["dram_bytes"] for k in rows)
            return dict(traffic=tot / n, traffic_unit="bytes/launch (dram read+write summed over ALL GEMM launches of one forward, ncu)", traffic_launches=n,
                        traffic_algorithmic_bytes_per_launch=d.get("gemm_algorithmic_bytes_per_launch"), traffic_source=str(p.relative_to(REPO)))
    return dict(traffic=None)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.samples, self.stop, self.index = [], threading.Event(), index
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=3)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in self.samples):
                reasons.append(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def inputs(h, image_index: int):
    x = h.randn(42 + 100 * image_index, (1, 4, 64, 64))
    cond = h.randn(43 + 100 * image_index, (1, 77, 768))
    uncond = h.randn(44, (1, 77, 768))
    return x, cond, uncond


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU backend on the box's host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sdb200 import Harness, FLAG_FLASH_ATTN
    from oracle.cpu_ref import load_cpu_oracle   # allowed here: this arm times the oracle itself
    h = Harness()
    variant = load_cpu_oracle(h)
    from oracle.cpu_ref import best_thread_count, usable_cores
    cores = best_thread_count(h)          # fastest ggml thread count on this host (128 logical CPUs may sit behind a much smaller quota)
    m = h.model("CPU", "sd15_unet", "f16", 0, 1234, cores)   # non-FA graph: the reference's default, and its fastest CPU path
    x, cond, uncond = inputs(h, 0)
    budget_s = float(os.environ.get("SDB200_REF_BUDGET_S", "150"))
    t0 = time.time()
    m.sample(x, cond, uncond, steps=1, cfg_scale=CFG_SCALE, eta=ETA)          # warm-up step (also sizes the run)
    first = time.time() - t0
    steps = max(1, min(args.steps, int(budget_s / max(first, 1e-3))))
    t0 = time.time()
    _, info = m.sample(x, cond, uncond, steps=steps, cfg_scale=CFG_SCALE, eta=ETA)
    dt = time.time() - t0
    m.close()
    val = steps / dt
    line = dict(impl="reference", metric="denoise_steps_per_s", value=val, unit="steps/s", n_gpus=0, steps=steps, warmup=1,
                ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16", data="synthetic",
                config=dict(workload=WORKLOAD, weights="F16 synthetic (seed 1234)", sampler="euler_a", cfg_scale=CFG_SCALE,
                            graph="reference default (MUL_MAT+SOFT_MAX attention)", forwards_per_step=2),
                cpu_baseline=dict(value=val, unit="steps/s", cores=cores, kind="reference",
                                  sample=f"{steps} full CFG denoise step(s) (2 UNet forwards each), ggml CPU backend variant '{variant}', {cores} threads; "
                                         f"steps capped from --steps {args.steps} to fit {budget_s:.0f} s"),
                e2e=dict(value=val, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.gpus
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sdb200 import Harness, FLAG_FLASH_ATTN
    h = Harness()
    devs = h.load_b200()                      # raises when libggml-b200.so is missing or no sm_100 device: no CPU fallback
    dev = f"B200_{local}"
    if dev not in devs:
        raise RuntimeError(f"{dev} not registered (devices: {devs})")
    m = h.model(dev, "sd15_unet", "f16", FLAG_FLASH_ATTN, 1234, 0)

    # ---- layouts.  One *step* is always the cond + uncond evaluation of one image plus the reference's host sampler math.
    #   "batched"   (default, any N): every GPU samples its own image; cond + uncond are ONE N = 2 forward per step (caller-side CFG
    #               batching); independent units, no data-path collective
    #   "serial"    (--cfg serial, N = 1): the reference's own order, two N = 1 forwards per step (stable-diffusion.cpp:2811-2836)
    #   "cfg-split" (--layout cfg-split, even N): ranks (2i, 2i+1) take cond / uncond of image i and all-gather eps over NCCL
    split_ok = world > 1 and world % 2 == 0
    main_layout = "cfg-split" if (split_ok and args.layout in ("auto", "cfg-split")) else ("serial" if (world == 1 and args.cfg == "serial") else "batched")
    alt_layout = None
    if not args.no_alt:
        if world == 1:
            alt_layout = "serial" if main_layout == "batched" else "batched"
        elif world % 2 == 0:
            alt_layout = "cfg-split" if main_layout == "batched" else "batched"

    x0, cond0, _ = inputs(h, 0)
    nodes, flops = m.dump_graph(None, x0, np.array([999.0], np.float32), cond0)
    assert abs(flops - FLOPS_PER_FORWARD) / FLOPS_PER_FORWARD < 0.01, f"graph FLOPs {flops:.4e} differ from SURVEY.md 8d"

    # CFG split: the exchange is DEVICE code over NVLink peer memory (kernels/peer.cu) -- each rank maps its partner's mailbox once (the two
    # 64-byte CUDA IPC handles travel over torch.distributed here, host plumbing) and from then on every forward ends with its eps
    # prediction stored into the partner's HBM by the output convolution's epilogue + a flag handshake, inside the replayed CUDA graph
    mailbox_ready = False

    def connect_mailbox():
        nonlocal mailbox_ready
        if mailbox_ready:
            return
        mine = torch.frombuffer(bytearray(m.mailbox_create(4 * 64 * 64 * 4)), dtype=torch.uint8).clone().cuda()
        handles = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(handles, mine)
        m.mailbox_connect(bytes(handles[rank ^ 1].cpu().numpy().tobytes()))
        mailbox_ready = True

    def nonlocal_reset():
        nonlocal mailbox_ready
        mailbox_ready = False

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(layout, steps, warmup, sample_clocks):
        """W untimed + exactly K timed steps in `layout`; device time = CUDA events around every graph_compute (+ the collective),
        max over ranks; returns the whole-job numbers."""
        if layout == "cfg-split":
            connect_mailbox()
            image, role, images = rank // 2, rank % 2, world // 2
        else:
            if mailbox_ready:
                m.mailbox_close()
                nonlocal_reset()
            image, role, images = rank, (2 if layout == "batched" else -1), world
        exchange = None
        x, cond, uncond = inputs(h, image)

        def run(k):
            return m.sample(x, cond, uncond, steps=k, cfg_scale=CFG_SCALE, eta=ETA, role=role, exchange=exchange)

        run(warmup)                                                # >= 3 untimed warm-up steps (plan caches, workspace, clocks)
        barrier()
        s0 = m.stats()
        clk = ClockSampler(local) if sample_clocks else None
        if clk:
            clk.__enter__()
        t0 = time.perf_counter()
        out, info = run(steps)                                     # EXACTLY K timed steps
        barrier()
        wall = time.perf_counter() - t0
        if clk:
            clk.__exit__(None, None, None)
        s1 = m.stats()
        dev_ms = s1["total_graph_ms"] - s0["total_graph_ms"]      # cfg-split: includes the in-graph exchange and the wait for the partner
        launches = s1["kernel_launches"] - s0["kernel_launches"]
        forwards = s1["graphs"] - s0["graphs"]
        if dist is not None:
            t = torch.tensor([wall, dev_ms, float(launches), float(forwards)], dtype=torch.float64, device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            wall, dev_ms, launches, forwards = float(tmax[0]), float(tmax[1]), float(t[2]), float(t[3])
        assert np.isfinite(out).all()
        return dict(layout=layout, images=images, value=images * steps / (dev_ms / 1e3), e2e=images * steps / wall, dev_ms=dev_ms, wall=wall,
                    launches=launches, forwards=forwards, clk=clk, s0=s0, s1=s1,
                    forwards_per_step=1 if layout != "serial" else 2)

    W = max(args.warmup, 3)
    main = timed(main_layout, args.steps, W, True)
    alt = timed(alt_layout, args.steps, W, False) if alt_layout else None
    if mailbox_ready:                 # the passes below are rank-local: no forward may wait for a partner any more
        m.mailbox_close()
        nonlocal_reset()
    clk, s0, s1 = main["clk"], main["s0"], main["s1"]
    wall, dev_ms, launches, forwards, images = main["wall"], main["dev_ms"], main["launches"], main["forwards"], main["images"]
    value, e2e = main["value"], main["e2e"]
    x, cond, uncond = inputs(h, rank if main_layout != "cfg-split" else rank // 2)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): per-launch CUDA events on the launching stream, separate pass
    roof = None
    cpu_base = None
    if rank == 0:
        pk = peaks()
        try:
            m.set_option("kernel_timing", 1)
            k0 = m.stats()
            # rank-0-local pass (no collective: the other ranks are not in this branch): 2 full CFG steps on this GPU alone
            m.sample(x, cond, uncond, steps=2, cfg_scale=CFG_SCALE, eta=ETA, role=-1 if main_layout == "serial" else 2, exchange=None)
            k1 = m.stats()
            m.set_option("kernel_timing", 0)
            gl = k1["tc_gemm_launches"] - k0["tc_gemm_launches"]
            gf = k1["tc_gemm_flops"] - k0["tc_gemm_flops"]
            gus = k1["tc_gemm_us"] - k0["tc_gemm_us"]
            if gus > 0:
                ach = gf / (gus * 1e-6) / 1e12
                roof = dict(bound="tensor", kernel="k_gemm_tc / k_gemm_tc2 (tcgen05.mma kind::f16, cta_group::1 and ::2, TMA, TMEM)", achieved=ach, peak=pk["bf16_sustained"],
                            unit="TFLOP/s", frac=ach / pk["bf16_sustained"], peak_source=pk["src"] + " (sustained: kernel timed inside a long step)",
                            launches=gl, flop_per_launch=gf / max(gl, 1), us_per_launch=gus / max(gl, 1),
                            share_of_step_device_time=(gus / 1e3) / max((k1["total_graph_ms"] - k0["total_graph_ms"]), 1e-9), **ncu_traffic())
        except Exception as e:   # older plugin without kernel timing
            roof = dict(bound="tensor", error=str(e))
        step_tflops = 2 * flops * images * args.steps / (dev_ms / 1e3) / 1e12 / max(1, world)
        vae = None
        if n == 1 and not args.no_vae:
            vae = vae_decode_leg(h, dev, pk)
            # the SDXL / Flux image size (BASELINE configs 3-4): 128x128 latent -> 1024x1024, algorithmic bytes 28.8 GB
            vae["at_1024"] = vae_decode_leg(h, dev, pk, latent=128, gb=28.8)
        if n == 1 and not args.no_cpu_baseline:
            cpu_base = cpu_baseline_leg(h)
    m.close()
    extra = {}
    if rank == 0 and n == 1:
        for name in [e for e in args.extra.split(",") if e and e != "none"]:
            extra[name] = extra_leg(h, dev, peaks(), name)
    if rank == 0:
        line = dict(metric="denoise_steps_per_s", value=value, unit="steps/s", n_gpus=n, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling=("strong" if (main_layout == "cfg-split" and world == 2) else "weak"),
                    vs_baseline=None, dtype="f16", data="synthetic",
                    config=dict(workload=WORKLOAD, denoiser="SD1.5 UNet graph built by the reference UNetModelRunner, synthetic F16 weights (seed 1234)",
                                latent="64x64x4", context="77x768", sampler="euler_a eta 1", cfg_scale=CFG_SCALE, forwards_per_step=main["forwards_per_step"], layout=main_layout,
                                graph="flash-attention variant (--diffusion-fa)", images=images,
                                parallelism=LAYOUT_TEXT[main_layout] + f"; {images} image(s) in flight on {world} GPU(s)",
                                l2="no explicit flush: every forward streams 1.72 GB of weights (> 126 MB L2)",
                                algorithmic_tflop_per_step=2 * flops / 1e12, step_tensor_frac_of_sustained_peak=step_tflops / peaks()["bf16_sustained"],
                                graph_nodes=nodes),
                    e2e=dict(value=e2e, unit="steps/s", ms_per_step=1e3 * wall / args.steps,
                             h2d_bytes_per_step=images * 2 * (4 * 64 * 64 * 4 + 77 * 768 * 4 + 4 + 8), d2h_bytes_per_step=images * 2 * 4 * 64 * 64 * 4),
                    gpu_launches=int(launches), forwards=int(forwards), clocks=clk.summary(), roofline=roof, cpu_baseline=cpu_base,
                    alt_layout=None if alt is None else dict(layout=alt["layout"], description=LAYOUT_TEXT[alt["layout"]], value=alt["value"], unit="steps/s",
                                                                             ms_per_step=alt["dev_ms"] / args.steps, e2e=alt["e2e"], images=alt["images"],
                                                                             forwards_per_step=alt["forwards_per_step"], gpu_launches=int(alt["launches"])),
                    vae_decode=vae, extra_workloads=extra or None,
                    layouts={main_layout: dict(value=value, e2e=e2e, ms_per_step=dev_ms / args.steps, e2e_ms_per_step=1e3 * wall / args.steps, is_value=True),
                             **({alt["layout"]: dict(value=alt["value"], e2e=alt["e2e"], ms_per_step=alt["dev_ms"] / args.steps,
                                                     e2e_ms_per_step=1e3 * alt["wall"] / args.steps, is_value=False)} if alt else {})},
                    host=dict(backend_graph_compute_ms_per_step=(s1["host_us"] - s0["host_us"]) / 1e3 / args.steps,
                              set_tensor_ms_per_step=(s1["host_set_us"] - s0["host_set_us"]) / 1e3 / args.steps,
                              get_tensor_incl_device_wait_ms_per_step=(s1["host_get_us"] - s0["host_get_us"]) / 1e3 / args.steps,
                              outside_backend_ms_per_step=(s1["host_outside_us"] - s0["host_outside_us"]) / 1e3 / args.steps,
                              e2e_minus_device_ms_per_step=1e3 * wall / args.steps - dev_ms / args.steps,
                              note="attribution at the plugin boundary: outside_backend = host time between two vtable calls (the reference's graph rebuild, "
                                   "gallocr, sampler math), set/get = inside the buffer vtable (get includes waiting for the device), "
                                   "backend_graph_compute = host side of graph_compute (signature, cudaGraphLaunch)"),
                    backend=dict(cuda_graph_replays=int(s1["cuda_graph_replays"] - s0["cuda_graph_replays"]),
                                 gemm_ref_launches=int(s1["gemm_ref_launches"] - s0["gemm_ref_launches"]),
                                 cta_pair_gemm_launches=int(s1["cta2_gemm_launches"] - s0["cta2_gemm_launches"]),
                                 unfused_attention=int(s1["unfused_attention"] - s0["unfused_attention"]),
                                 fused_nodes=int(s1["fused_nodes"] - s0["fused_nodes"]), implicit_convs=int(s1["implicit_convs"] - s0["implicit_convs"]),
                                 fused_attn_launches=int(s1["fused_attn_launches"] - s0["fused_attn_launches"]),
                                 tc_gemm_launches=int(s1["tc_gemm_launches"] - s0["tc_gemm_launches"])))
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def vae_decode_leg(h, dev, pk, latent=64, gb=7.27):
    """Second half of BASELINE.json's metric: AutoEncoderKL decode 64x64x4 -> 512x512x3 (reference graph, synthetic F16 weights).
    HBM roofline per the north star: algorithmic bytes 7.27 GB / device time vs measured copy bandwidth;
    the tensor fraction (2.515 TFLOP) is reported beside it because the fused graph crosses the ridge."""
    m = h.model(dev, "vae_decoder", "f16", 0, 1234, 0)
    z = h.randn(45, (1, 4, latent, latent))
    nodes, flops = m.dump_graph(None, z)
    for _ in range(3):
        m.forward(z)
    dev_ms, wall_ms = [], []
    for _ in range(5):
        s0 = m.stats()
        t0 = time.perf_counter()
        out, _ = m.forward(z)
        wall_ms.append((time.perf_counter() - t0) * 1e3)
        s1 = m.stats()
        dev_ms.append(s1["total_graph_ms"] - s0["total_graph_ms"])
    m.close()
    d = statistics.median(dev_ms)
    return dict(metric="vae_decode_ms", value=d, unit="ms", e2e_ms=statistics.median(wall_ms),
                workload=f"AutoEncoderKL decode {latent}x{latent}x4 -> {8 * latent}x{8 * latent}x3, F16 conv weights",
                algorithmic_gb=gb, hbm_gbs=gb / (d / 1e3), hbm_frac=gb / (d / 1e3) / pk["hbm"], algorithmic_tflop=flops / 1e12,
                tensor_tflops=flops / 1e12 / (d / 1e3), tensor_frac=flops / 1e12 / (d / 1e3) / pk["bf16_sustained"], graph_nodes=nodes,
                finite=bool(np.isfinite(out).all()))


EXTRA = {
    # BASELINE.json configs[2] / configs[3] shapes at one GPU (the multi-GPU layouts of those configs shard these same forwards)
    "sdxl": dict(arch="sdxl_unet", wtype="bf16", flags=1, x=(1, 4, 128, 128), ctx=(1, 77, 2048), y=(1, 2816), t=999.0, per_step=2,
                 workload="SDXL UNet forward, 128x128x4 latent (1024x1024), BF16 linears / F16 convs, flash-attention graph"),
    "flux": dict(arch="flux_schnell", wtype="bf16", flags=1, x=(1, 16, 128, 128), ctx=(1, 256, 4096), y=(1, 768), t=1.0, per_step=1,
                 workload="FLUX.1-schnell MMDiT forward, 4096 image + 256 text tokens (1024x1024), BF16 weights"),
}


def extra_leg(h, dev, pk, name):
    """Forward time of a larger north-star config on ONE GPU: device ms from CUDA events around graph_compute (weights resident),
    algorithmic FLOPs recomputed from the live ggml graph, fraction of the measured sustained bf16 peak."""
    e = EXTRA[name]
    t0 = time.time()
    m = h.model(dev, e["arch"], e["wtype"], e["flags"], 1234, 0)
    create_s = time.time() - t0
    x = h.randn(42, e["x"]); ctx = h.randn(43, e["ctx"]); y = h.randn(44, e["y"]); t = np.array([e["t"]], np.float32)
    nodes, flops = m.dump_graph(None, x, t, ctx, y)
    for _ in range(3):
        out, _ = m.forward(x, t, ctx, y)
    dev_ms, wall_ms = [], []
    for _ in range(5):
        s0 = m.stats()
        t0 = time.perf_counter()
        out, _ = m.forward(x, t, ctx, y)
        wall_ms.append((time.perf_counter() - t0) * 1e3)
        s1 = m.stats()
        dev_ms.append(s1["total_graph_ms"] - s0["total_graph_ms"])
    launches = (s1["kernel_launches"] - s0["kernel_launches"])
    m.close()
    d = statistics.median(dev_ms)
    return dict(workload=e["workload"], forward_ms=d, forward_e2e_ms=statistics.median(wall_ms), steps_per_s=1e3 / (d * e["per_step"]),
                forwards_per_step=e["per_step"], algorithmic_tflop=flops / 1e12, tensor_tflops=flops / 1e12 / (d / 1e3),
                tensor_frac=flops / 1e12 / (d / 1e3) / pk["bf16_sustained"], graph_nodes=nodes, launches_per_forward=int(launches),
                finite=bool(np.isfinite(out).all()), model_create_s=create_s)


def cpu_baseline_leg(h):
    """Bounded CPU sample: ONE full CFG denoise step (2 UNet forwards) on the reference CPU backend, all host threads."""
    from oracle.cpu_ref import load_cpu_oracle, best_thread_count
    variant = load_cpu_oracle(h)
    cores = best_thread_count(h)
    m = h.model("CPU", "sd15_unet", "f16", 0, 1234, cores)
    x, cond, uncond = inputs(h, 0)
    t0 = time.time()
    m.sample(x, cond, uncond, steps=1, cfg_scale=CFG_SCALE, eta=ETA)
    dt = time.time() - t0
    m.close()
    return dict(value=1.0 / dt, unit="steps/s", cores=cores, kind="reference",
                sample=f"1 full CFG denoise step (2 SD1.5 UNet forwards) on the reference ggml CPU backend ('{variant}' variant, {cores} threads), {dt:.1f} s")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vae", action="store_true")
    ap.add_argument("--cfg", default="batched", choices=["batched", "serial"],
                    help="N = 1: cond + uncond as ONE N = 2 forward per step (default; caller-side CFG batching, SURVEY.md 8f-1) or the "
                         "reference's two serial forwards; the other one is measured too and reported under alt_layout")
    ap.add_argument("--layout", default="auto", choices=["auto", "batched", "cfg-split"],
                    help="N > 1: 'cfg-split' (auto at even N: the north star's layout) splits the CFG batch of one image over a GPU pair with a "
                         "device-side exchange of eps per step; 'batched' runs independent images per GPU with batched CFG (no collective); the "
                         "other one is reported under alt_layout")
    ap.add_argument("--no-alt", action="store_true", help="skip the alt_layout measurement")
    ap.add_argument("--extra", default="sdxl,flux", help="comma list of additional single-GPU forward timings (N = 1 only): sdxl,flux; 'none' disables")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
