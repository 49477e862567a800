"""Generates the committed golden fixtures under tests/golden/.  Run HERE (the container that has
/root/reference and the compiled oracle); the fixtures are what travels to the GPU box.

  ggml_test_conv2d.json    the reference's own known-answer vectors, transcribed from
                           ggml/tests/test-conv2d.cpp:237-360 (expected_conv2d f32[480], expected_im2col u16[480];
                           KW=KH=3, IC=OC=10, IW=8, IH=6, N=1, kernel == 2.5, image == 1.5, s=1, p=1, d=1)
  schedule_sd15.json       sigmas / timesteps of the reference scheduler (the reference's denoiser.hpp compiled into
                           host/_ref/libsd_harness.so) for 1..50 steps, as hex bit patterns: the bit-exactness pin
  philox_seed42.json       first 64 values of the reference Philox randn stream (core/rng_philox.hpp)
  cpu_ops.npz              seeded inputs and the reference CPU backend's outputs (oracle/_ref) for each hot-path op
  cpu_models.npz           reference CPU backend outputs of whole synthetic-weight models (unet_tiny fa/non-fa,
                           vae_decoder on an 8x8 latent, scheduler-driven 3-step sample)
  cpu_wan_vae.npz          (`make_golden.py wan_vae`) Wan causal-3D VAE decoder, one latent frame
  truth_f64.npz            (`make_golden.py truth`) float64 evaluation (oracle/graph_f64.py) of the graphs the reference builds for unet_tiny and for
                           the SD1.5 UNet at 64x64 (default attention graph): the arbiter of the whole-model parity tests, plus the
                           reference CPU backend's distance to it (cpu_rel_*)
  cpu_models_dit.npz       (`make_golden.py dit`) the same for the larger architectures of SURVEY.md 8a: SD1.5 UNet 64x64 (default graph),
                           SDXL UNet 32x32, flux_tiny (bf16), SD3-medium MMDiT 32x32 (f16), Wan2.1-1.3B DiT (q8_0, 3x16x16 latent)
"""
import json
import re
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
sys.path.insert(0, str(REPO / "stable-diffusion.cpp_b200"))
sys.path.insert(0, str(REPO))
REF = Path("/root/reference")


def conv2d_vectors():
    src = (REF / "ggml/tests/test-conv2d.cpp").read_text()
    m1 = re.search(r"float expected_conv2d \[n_conv2d_test\] = \{(.*?)\};", src, re.S)
    m2 = re.search(r"uint16_t expected_im2col\[n_conv2d_test\] = \{(.*?)\};", src, re.S)
    conv = [float(v.strip().rstrip("f")) for v in m1.group(1).replace("\n", " ").split(",") if v.strip()]
    im2 = [int(v.strip()) for v in m2.group(1).replace("\n", " ").split(",") if v.strip()]
    assert len(conv) == 480 and len(im2) == 480
    (HERE / "ggml_test_conv2d.json").write_text(json.dumps(dict(
        source="ggml/tests/test-conv2d.cpp:237-360", KW=3, KH=3, IC=10, OC=10, IW=8, IH=6, N=1, kernel_value=2.5, image_value=1.5,
        s=1, p=1, d=1, expected_conv2d=conv, expected_im2col_u16=im2)))


def main():
    from sdb200 import Harness, FLAG_FLASH_ATTN
    from oracle.cpu_ref import load_cpu_oracle
    conv2d_vectors()
    h = Harness()
    load_cpu_oracle(h)

    sched = {}
    for steps in (1, 2, 4, 10, 20, 30, 50):
        s, t = h.schedule(steps)
        sched[str(steps)] = dict(sigmas_hex=[f"{v:08x}" for v in s.view(np.uint32)], timesteps_hex=[f"{v:08x}" for v in t.view(np.uint32)])
    (HERE / "schedule_sd15.json").write_text(json.dumps(sched))
    r = h.randn(42, (64,))
    (HERE / "philox_seed42.json").write_text(json.dumps(dict(seed=42, n=64, values_hex=[f"{v:08x}" for v in r.view(np.uint32)])))

    rng = np.random.default_rng(1234)
    f = lambda *shape: rng.standard_normal(shape).astype(np.float32)
    ops = {}
    x = f(1, 64, 8, 8); w = 1 + 0.1 * f(1, 64, 1, 1); b = 0.1 * f(1, 64, 1, 1)
    ops["gn_x"], ops["gn_w"], ops["gn_b"] = x, w, b
    ops["gn_y"] = h.run_op("CPU", "group_norm", [x], ip=[32, 0], fp=[1e-6])
    ops["gn_affine_silu_y"] = h.run_op("CPU", "group_norm", [x, w, b], ip=[32, 1], fp=[1e-6])
    x = f(1, 1, 6, 320)
    ops["ln_x"], ops["ln_y"], ops["rms_y"] = x, h.run_op("CPU", "norm", [x], fp=[1e-5]), h.run_op("CPU", "rms_norm", [x], fp=[1e-6])
    x = f(1, 2, 5, 77)
    ops["sm_x"], ops["sm_y"] = x, h.run_op("CPU", "soft_max", [x], fp=[0.125, 0.0])
    wm, xm = f(48, 72) / 8, f(10, 72)
    ops["mm_w"], ops["mm_x"] = wm, xm
    ops["mm_f32_y"] = h.run_op("CPU", "mul_mat", [wm, xm], ["f32", "f32"])
    ops["mm_f16_y"] = h.run_op("CPU", "mul_mat", [wm, xm], ["f16", "f32"])
    ops["mm_bf16_y"] = h.run_op("CPU", "mul_mat", [wm, xm], ["bf16", "f32"])
    cw, cx, cb = f(16, 8, 3, 3) / 8, f(1, 8, 12, 12), 0.1 * f(1, 16, 1, 1)
    ops["conv_w"], ops["conv_x"], ops["conv_b"] = cw, cx, cb
    ops["conv_y"] = h.run_op("CPU", "conv_2d", [cw, cx, cb], ["f16", "f32", "f32"], ip=[1, 1, 1, 1, 1, 1])
    ops["conv_s2_y"] = h.run_op("CPU", "conv_2d", [cw, cx, cb], ["f16", "f32", "f32"], ip=[2, 2, 1, 1, 1, 1])
    q, k, v = f(1, 2, 20, 40), f(1, 2, 33, 40), f(1, 2, 33, 40)
    ops["fa_q"], ops["fa_k"], ops["fa_v"] = q, k, v
    ops["fa_y"] = h.run_op("CPU", "flash_attn", [q, k, v], ["f32", "f16", "f16"], fp=[40 ** -0.5])
    t = np.array([999.0, 500.5, 1.0], np.float32)
    ops["ts_t"], ops["ts_y"] = t, h.run_op("CPU", "timestep_embedding", [t], ip=[320, 10000])
    x = f(1, 3, 4, 5)
    ops["up_x"], ops["up_y"] = x, h.run_op("CPU", "upscale", [x], ip=[2, 0])
    x = np.linspace(-6, 6, 97, dtype=np.float32).reshape(1, 1, 1, 97)
    ops["act_x"], ops["silu_y"], ops["gelu_y"] = x, h.run_op("CPU", "unary", [x], ip=[10]), h.run_op("CPU", "unary", [x], ip=[8])
    np.savez_compressed(HERE / "cpu_ops.npz", **ops)

    models = {}
    x = h.randn(42, (1, 4, 16, 16)); ctx = h.randn(43, (1, 77, 768)); tt = np.array([999.0], np.float32)
    for fa in (0, 1):
        m = h.model("CPU", "unet_tiny", "f16", fa, 1234, 8)
        models[f"unet_tiny_fa{fa}"], _ = m.forward(x, tt, ctx)
        if fa == 0:
            unc = h.randn(44, (1, 77, 768))
            out, info = m.sample(x, ctx, unc, steps=3, cfg_scale=7.0, eta=1.0, method="euler_a", sampler_seed=42)
            models["unet_tiny_sample3"] = out
        m.close()
    m = h.model("CPU", "vae_decoder", "f16", 0, 1234, 8)
    z = h.randn(45, (1, 4, 8, 8))
    models["vae_decoder_8x8"], _ = m.forward(z)
    m.close()
    np.savez_compressed(HERE / "cpu_models.npz", **models)
    print("golden fixtures written to", HERE)


DIT_CASES = {
    # key: (arch, wtype, flags, x shape, context shape, y shape or None, timestep)
    "sd15_unet_fa0": ("sd15_unet", "f16", 0, (1, 4, 64, 64), (1, 77, 768), None, 999.0),
    "sdxl_unet_32": ("sdxl_unet", "f16", 0, (1, 4, 32, 32), (1, 77, 2048), (1, 2816), 999.0),
    "flux_tiny": ("flux_tiny", "bf16", 1, (1, 16, 32, 32), (1, 64, 4096), (1, 768), 1.0),
    "mmdit_sd3": ("mmdit_sd3", "f16", 1, (1, 16, 32, 32), (1, 154, 4096), (1, 2048), 500.0),
    "wan_1_3b": ("wan_1_3b", "q8_0", 1, (16, 3, 16, 16), (1, 512, 4096), None, 500.0),
}


def dit_inputs(h, key):
    arch, wtype, flags, xs, cs, ys, t = DIT_CASES[key]
    return h.randn(42, xs), np.array([t], np.float32), h.randn(43, cs), (h.randn(44, ys) if ys else None)


def main_dit():
    from sdb200 import Harness
    from oracle.cpu_ref import load_cpu_oracle
    h = Harness()
    variant = load_cpu_oracle(h)
    out = {}
    for key, (arch, wtype, flags, *_rest) in DIT_CASES.items():
        m = h.model("CPU", arch, wtype, flags, 1234, 8)
        x, t, ctx, y = dit_inputs(h, key)
        out[key], _ = m.forward(x, t, ctx, y)
        m.close()
        print(key, out[key].shape, float(out[key].std()), flush=True)
    np.savez_compressed(HERE / "cpu_models_dit.npz", **out)
    print("written with CPU variant", variant)


def main_wan_vae():
    """cpu_wan_vae.npz: Wan causal-3D VAE decoder, ONE latent frame [16, 1, 8, 8] -> [3, 1, 64, 64] (with more frames the reference's
    single-graph decode path yields NaN from the second frame on, on its own CPU backend -- wan_vae.hpp:1381 calls its chunked twin
    "weird" too -- so only the first frame can be pinned)."""
    from sdb200 import Harness
    from oracle.cpu_ref import load_cpu_oracle
    h = Harness()
    load_cpu_oracle(h)
    m = h.model("CPU", "wan_vae_decoder", "f16", 0, 1234, 8)
    out, _ = m.forward(h.randn(45, (16, 1, 8, 8)))
    m.close()
    assert np.isfinite(out).all()
    np.savez_compressed(HERE / "cpu_wan_vae.npz", wan_vae_1frame=out)
    print("wan_vae_1frame", out.shape, float(out.std()))


def main_truth():
    """truth_f64.npz: what the graph evaluates to in exact (float64, no intermediate rounding) arithmetic, and how far the reference CPU
    backend is from it.  SD1.5 at 64x64 takes ~2 min of numpy and ~3 GB."""
    import tempfile
    from sdb200 import Harness
    from oracle.cpu_ref import load_cpu_oracle
    from oracle.graph_f64 import evaluate
    h = Harness()
    load_cpu_oracle(h)
    out = {}
    for key, arch, shape in (("unet_tiny", "unet_tiny", (1, 4, 16, 16)), ("sd15_unet", "sd15_unet", (1, 4, 64, 64))):
        x = h.randn(42, shape); ctx = h.randn(43, (1, 77, 768)); t = np.array([999.0], np.float32)
        m = h.model("CPU", arch, "f16", 0, 1234, 8)
        cpu, _ = m.forward(x, t, ctx)
        with tempfile.TemporaryDirectory() as d:
            m.export_graph(Path(d) / "g", x, t, ctx)
            m.close()
            truth = evaluate(Path(d) / "g").reshape(cpu.shape)
        out[key] = truth
        out["cpu_rel_" + key] = np.float64(np.linalg.norm(cpu.astype(np.float64) - truth) / np.linalg.norm(truth))
        print(key, truth.shape, "cpu vs truth rel_l2", float(out["cpu_rel_" + key]), flush=True)
    np.savez_compressed(HERE / "truth_f64.npz", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "truth":
        main_truth()
    elif len(sys.argv) > 1 and sys.argv[1] == "dit":
        main_dit()
    elif len(sys.argv) > 1 and sys.argv[1] == "wan_vae":
        main_wan_vae()
    else:
        main()
