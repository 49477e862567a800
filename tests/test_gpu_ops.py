"""GPU parity tests proper: every hot-path op, through the C-ABI (ggml vtable -> libggml-b200.so), against the
reference CPU backend on identical seeded inputs, at the shapes the SD1.5 / SDXL / VAE / Flux graphs use.
Tolerances: test-backend-ops' NMSE limits (ggml/tests/test-backend-ops.cpp: 1e-7 default, 5e-4 MUL_MAT / FLASH_ATTN_EXT /
conv) restated as relative-L2 (sqrt of NMSE)."""
import numpy as np
import pytest

from oracle import ops_ref as R

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def both(b200, op, inputs, itypes=None, ip=(), fp=()):
    h, dev = b200
    return h.run_op(dev, op, inputs, itypes, ip, fp), h.run_op("CPU", op, inputs, itypes, ip, fp)


RNG = np.random.default_rng(2024)
f = lambda *s: RNG.standard_normal(s).astype(np.float32)


@pytest.mark.parametrize("C,H,W", [(320, 64, 64), (640, 32, 32), (1280, 16, 16), (1280, 8, 8), (2560, 8, 8), (128, 96, 96), (30, 5, 7)])
def test_group_norm(b200, C, H, W):
    x = f(1, C, H, W) * 3 + 0.5
    g, c = both(b200, "group_norm", [x], ip=[32 if C % 32 == 0 else 6, 0], fp=[1e-6])
    assert rel(g, c) < 3e-4 ** 0.5 * 1e-2      # ~3e-6


def test_group_norm_affine_silu_chain(b200):
    x, w, b = f(2, 320, 32, 32), 1 + 0.1 * f(1, 320, 1, 1), 0.1 * f(1, 320, 1, 1)
    g, c = both(b200, "group_norm", [x, w, b], ip=[32, 1], fp=[1e-6])
    assert rel(g, c) < 5e-6


@pytest.mark.parametrize("shape", [(1, 1, 4096, 320), (1, 1, 77, 768), (1, 2, 256, 1280), (1, 1, 4352, 3072)])
def test_layer_and_rms_norm(b200, shape):
    x = f(*shape) * 2 + 0.3
    for op, eps in (("norm", 1e-5), ("rms_norm", 1e-6)):
        g, c = both(b200, op, [x], fp=[eps])
        assert rel(g, c) < 5e-6, op


@pytest.mark.parametrize("shape", [(1, 8, 64, 77), (1, 8, 256, 256), (1, 2, 128, 4096)])
def test_soft_max(b200, shape):
    x = f(*shape) * 4
    g, c = both(b200, "soft_max", [x], fp=[0.158, 0.0])
    assert rel(g, c) < 5e-6


@pytest.mark.parametrize("wt", ["f16", "f32", "bf16"])
@pytest.mark.parametrize("M,N,K", [(320, 4096, 320), (2560, 1024, 640), (1280, 77, 768), (320, 64, 2880), (1280, 256, 11520), (64, 33, 40), (9, 5, 36),
                                   (1280, 1, 320), (320, 1, 1280), (1283, 3, 2816), (640, 4, 1280), (1280, 5, 1280)])   # N <= 4: weight-streaming GEMV
def test_mul_mat(b200, wt, M, N, K):
    w, x = f(M, K) / np.sqrt(K), f(N, K)
    g, c = both(b200, "mul_mat", [w, x], [wt, "f32"])
    # f32 x f32 runs as 3xTF32 (hi / lo operand split, three tensor-core passes, f32 accumulation): f32-class accuracy like the oracle's
    # f32 dot; f16 / bf16 match the oracle's rounding of the activation to the weight type
    assert rel(g, c) < (5e-6 if wt == "f32" else 2e-4), f"rel {rel(g, c):.2e}"


def test_mul_mat_f32_single_pass_tf32_is_the_opt_out(b200):
    """Option precise_f32 = 0 (GGML_B200_PRECISE_F32=0) keeps the one-pass TF32 contraction (10 mantissa bits per operand): 2e-3 class."""
    import os
    h, dev = b200
    w, x = f(640, 768) / np.sqrt(768), f(256, 768)
    exact = (x.astype(np.float64) @ w.astype(np.float64).T).astype(np.float32)
    g = h.run_op(dev, "mul_mat", [w, x], ["f32", "f32"])
    assert rel(g, exact) < 2e-6, f"3xTF32 vs f64: {rel(g, exact):.2e}"


@pytest.mark.parametrize("M,N,K", [(1536, 192, 1536), (8960, 704, 1536), (1536, 512, 4096), (320, 3, 256), (96, 40, 64)])
def test_mul_mat_q8_0_weights(b200, M, N, K):
    """BASELINE config 5 data format: Q8_0 weight blocks (ggml-common.h:251-255).  The oracle also quantises the ACTIVATION rows to Q8_0
    (ggml-cpu.c:1480-1510, vec_dot q8_0 x q8_0); so does this backend (the activation operand is written as d * q per 32-value block by
    its pack kernel), so the two contract the same values: gate at f16-rounding level against the oracle, and check the exact arithmetic
    against the f64 product of the dequantised operands."""
    w, x = f(M, K) / np.sqrt(K), f(N, K)
    g, c = both(b200, "mul_mat", [w, x], ["q8_0", "f32"])
    assert rel(g, c) < 1e-3, f"vs the CPU oracle (q8_0 x q8_0): {rel(g, c):.2e}"
    wq = R.dequant_q8_0(R.quant_q8_0(w))
    xq = R.dequant_q8_0(R.quant_q8_0(x))
    exact = xq.astype(np.float64) @ wq.astype(np.float64).T
    assert rel(g, exact.astype(np.float32)) < 1e-3, f"vs the product of the dequantised operands {rel(g, exact):.2e}"


@pytest.mark.parametrize("H,L,d,rms,f16out", [(24, 4352, 128, 1, 0), (24, 4352, 128, 0, 1), (3, 77, 64, 1, 1), (12, 200, 128, 0, 0)])
def test_rope_chain(b200, H, L, d, rms, f16out):
    """Rope::apply_rope (8 working nodes: two CONTs, two REPEATs, CONT(pe), two MULs, ADD) [+ QKNorm in front, + F16 cast behind] runs as one
    kernel; the products are rounded separately like the unfused nodes, so the result matches the CPU to f32 rounding."""
    x = f(1, L, H, d)
    ang = np.random.default_rng(5).uniform(-3.1, 3.1, (L, d // 2)).astype(np.float32)
    pe = np.stack([np.stack([np.cos(ang), -np.sin(ang)], -1), np.stack([np.sin(ang), np.cos(ang)], -1)], -2).astype(np.float32)   # [L, d/2, 2, 2]
    w = (1 + 0.1 * f(d)).astype(np.float32) if rms else None
    g, c = both(b200, "rope", [x, pe, w], ip=[f16out], fp=[1e-6])
    assert rel(g, c) < (1e-3 if f16out else 3e-6), f"rel {rel(g, c):.2e}"
    ref = R.rope_interleaved(R.rms_norm(x, 1e-6) * w if rms else x, pe)
    assert rel(g.reshape(ref.shape), ref) < (1e-3 if f16out else 3e-6), f"vs restatement {rel(g.reshape(ref.shape), ref):.2e}"


def test_conv_3d_patch_embedding(b200):
    """Wan patch embedding (wan.hpp): Conv3d kernel (1,2,2) stride (1,2,2) -> IM2COL_3D + MUL_MAT + permute."""
    IC, OC = 16, 1536
    w, x = f(OC * IC, 1, 2, 2) / 8, f(IC, 3, 16, 16)
    g, c = both(b200, "conv_3d", [w, x], ["f16", "f32"], ip=[IC, 2, 2, 1, 0, 0, 0, 1, 1, 1])
    assert g.shape == c.shape and rel(g, c) < 2e-4, f"rel {rel(g, c):.2e}"
    # causal-style 3x3x3 with padding
    w, x = f(8 * 4, 3, 3, 3) / 10, f(4, 5, 9, 7)
    g, c = both(b200, "conv_3d", [w, x], ["f16", "f32"], ip=[4, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    assert g.shape == c.shape and rel(g, c) < 2e-4, f"rel {rel(g, c):.2e}"


def test_mul_mat_batched_broadcast(b200):
    w, x = f(1, 2, 64, 80) / 9, f(1, 8, 100, 80)      # ne02 = 2 broadcast over ne12 = 8 (GQA-style)
    g, c = both(b200, "mul_mat", [w, x], ["f16", "f32"])
    assert rel(g, c) < 2e-4


@pytest.mark.parametrize("IC,OC,H,s", [(4, 320, 64, 1), (320, 320, 64, 1), (320, 640, 32, 2), (1280, 1280, 8, 1), (512, 3, 32, 1), (128, 128, 48, 1)])
def test_conv_2d(b200, IC, OC, H, s):
    w, x, b = f(OC, IC, 3, 3) / np.sqrt(9 * IC), f(1, IC, H, H), 0.1 * f(1, OC, 1, 1)
    g, c = both(b200, "conv_2d", [w, x, b], ["f16", "f32", "f32"], ip=[s, s, 1, 1, 1, 1])
    assert rel(g, c) < 2e-4, f"rel {rel(g, c):.2e}"


@pytest.mark.parametrize("N,IC,OC,H,W", [(2, 320, 320, 32, 64), (1, 128, 128, 48, 24), (2, 1280, 1280, 16, 16), (1, 640, 320, 32, 32), (1, 64, 128, 16, 8),
                                         (1, 1920, 640, 32, 32), (1, 256, 256, 64, 64)])
def test_conv_2d_3x3_patch_tiles(b200, N, IC, OC, H, W):
    """3x3 / stride 1 / pad 1 convolutions inside the halo-reuse envelope of the CTA-pair kernel (W % 8 == 0, H % 16 == 0: one image box in
    shared memory serves 3 or 9 taps, 16 x 8 pixel patches): batch > 1, an odd number of patches, one patch pair per image, split-K
    (few patches, long K), 3 and 9 taps per ring stage."""
    w, x, b = f(OC, IC, 3, 3) / np.sqrt(9 * IC), f(N, IC, H, W), 0.1 * f(1, OC, 1, 1)
    g, c = both(b200, "conv_2d", [w, x, b], ["f16", "f32", "f32"], ip=[1, 1, 1, 1, 1, 1])
    assert g.shape == c.shape and rel(g, c) < 2e-4, f"rel {rel(g, c):.2e}"


def test_conv_1x1(b200):
    w, x = f(640, 320, 1, 1) / 18, f(1, 320, 32, 32)
    g, c = both(b200, "conv_2d", [w, x], ["f16", "f32"], ip=[1, 1, 0, 0, 1, 1])
    assert rel(g, c) < 2e-4


@pytest.mark.parametrize("H,d,Lq,Lk", [(8, 40, 1024, 1024), (8, 80, 256, 256), (8, 160, 64, 64), (8, 160, 256, 77), (8, 40, 512, 77), (10, 64, 300, 300),
                                       (4, 128, 333, 589), (1, 512, 256, 256)])
def test_flash_attn_ext(b200, H, d, Lq, Lk):
    q, k, v = f(1, H, Lq, d), f(1, H, Lk, d), f(1, H, Lk, d)
    g, c = both(b200, "flash_attn", [q, k, v], ["f32", "f16", "f16"], fp=[d ** -0.5])
    ref = R.flash_attn_ext(q, k, v, None, d ** -0.5)            # f64 restatement: the arbiter when the CPU path rounds P.V to f16
    assert rel(g, ref) < 2e-3, f"vs restatement {rel(g, ref):.2e}"
    assert rel(g, c) < 2.3e-2, f"vs CPU oracle {rel(g, c):.2e}"  # sqrt(5e-4): test-backend-ops' FLASH_ATTN_EXT limit


def test_flash_attn_ext_left_padded_mask(b200):
    """A mask whose LEADING key tiles are fully -inf and whose open keys carry a large negative bias: the running maximum must stay
    -inf until the first open key (a reference of 0 would flush the later probabilities to f16 zeros and write zero rows)."""
    H, d, Lq, Lk = 4, 64, 130, 320
    q, k, v = f(1, H, Lq, d), f(1, H, Lk, d), f(1, H, Lk, d)
    mask = np.full((1, 1, Lq, Lk), -30.0, np.float32)
    mask[..., :160] = -np.inf
    mask[:, :, 7, 160:200] = -np.inf                              # one row opens even later
    g, c = both(b200, "flash_attn", [q, k, v, mask], ["f32", "f16", "f16", "f16"], fp=[d ** -0.5])
    ref = R.flash_attn_ext(q, k, v, mask, d ** -0.5)
    assert np.abs(ref).max() > 0.1
    assert rel(g, ref) < 2e-3, f"vs restatement {rel(g, ref):.2e}"
    assert rel(g, c) < 2.3e-2, f"vs CPU oracle {rel(g, c):.2e}"


@pytest.mark.parametrize("fa", [0, 1])
def test_attention_wrapper_both_graph_variants(b200, fa):
    """ggml_ext_attention_ext (ggml_extend.hpp:1349): reshape/permute/cont + (FLASH_ATTN_EXT | MUL_MAT+SOFT_MAX+MUL_MAT)."""
    q, k, v = f(1, 256, 640), f(1, 77, 640), f(1, 77, 640)
    g, c = both(b200, "attention", [q, k, v], ip=[8, fa])
    assert rel(g, c) < (2.3e-2 if fa else 3e-3)


def test_small_data_movement_ops(b200):
    x = f(1, 3, 17, 19)
    for op, ip in (("upscale", [2, 0]), ("cont_permute", [1, 2, 0, 3]), ("cont_permute", [0, 2, 1, 3]), ("cpy", [1])):
        g, c = both(b200, op, [x], ip=ip)
        assert np.array_equal(g, c), op
    a, b_ = f(1, 320, 8, 8), f(1, 640, 8, 8)
    g, c = both(b200, "concat", [a, b_], ip=[2])
    assert np.array_equal(g, c)
    # dim-0 concat (single-stream Flux block) and the broadcast adds of ResBlocks: per-channel vector shared by / distinct per image
    g, c = both(b200, "concat", [f(1, 1, 50, 3072), f(1, 1, 50, 12288)], ip=[0])
    assert np.array_equal(g, c)
    for bshape in ((1, 320, 1, 1), (2, 320, 1, 1)):
        g, c = both(b200, "add", [f(2, 320, 16, 16), f(*bshape)])
        assert np.array_equal(g, c), bshape
        g, c = both(b200, "mul", [f(2, 320, 16, 16), f(*bshape)])
        assert np.array_equal(g, c), bshape
    t = np.array([999.0, 500.5, 1.0], np.float32)
    g, c = both(b200, "timestep_embedding", [t], ip=[320, 10000])
    assert np.abs(g - c).max() < 2e-4                        # cosf/sinf at ~1e3 rad: libm vs CUDA differ in the last ulps of the argument reduction
    for uop in (10, 8):                                      # SILU, GELU (CPU GELU goes through an f16 table)
        g, c = both(b200, "unary", [f(1, 1, 64, 320)], ip=[uop])
        assert rel(g, c) < (1e-6 if uop == 10 else 2e-3)


@pytest.mark.parametrize("C,OC,H,silu", [(320, 320, 64, 1), (640, 1280, 16, 1), (1280, 1280, 8, 1), (320, 320, 32, 0), (128, 256, 128, 1)])
def test_resblock_prologue_plus_conv(b200, C, OC, H, silu):
    """GROUP_NORM -> MUL -> ADD -> SILU -> IM2COL -> MUL_MAT -> RESHAPE -> PERMUTE -> CONT: one fused implicit-GEMM launch chain."""
    x, gw, gb = f(1, C, H, H) * 2 + 0.3, 1 + 0.1 * f(1, C, 1, 1), 0.1 * f(1, C, 1, 1)
    w = f(OC, C, 3, 3) / np.sqrt(9 * C)
    g, c = both(b200, "gn_silu_conv", [x, gw, gb, w], ["f32", "f32", "f32", "f16"], ip=[32, silu, 1], fp=[1e-6])
    assert rel(g, c) < 3e-4, f"rel {rel(g, c):.2e}"


@pytest.mark.parametrize("C,OC,H", [(1280, 1280, 8), (640, 640, 32), (256, 256, 128)])
def test_upsample_conv(b200, C, OC, H):
    x, w, b = f(1, C, H, H), f(OC, C, 3, 3) / np.sqrt(9 * C), 0.1 * f(1, OC, 1, 1)
    g, c = both(b200, "upscale_conv", [x, w, b], ["f32", "f16", "f32"])
    assert rel(g, c) < 3e-4, f"rel {rel(g, c):.2e}"


def test_implicit_conv_matches_unfused_path(b200):
    """Same conv with fusion off (IM2COL + GEMM as separate nodes) and on (implicit GEMM): must agree to f32 summation noise."""
    import os
    h, dev = b200
    w, x, b = f(640, 320, 3, 3) / 54, f(1, 320, 32, 32), 0.1 * f(1, 640, 1, 1)
    fused = h.run_op(dev, "conv_2d", [w, x, b], ["f16", "f32", "f32"], ip=[1, 1, 1, 1, 1, 1])
    os.environ["GGML_B200_FUSION"] = "0"
    try:
        plain = h.run_op(dev, "conv_2d", [w, x, b], ["f16", "f32", "f32"], ip=[1, 1, 1, 1, 1, 1])
    finally:
        del os.environ["GGML_B200_FUSION"]
    assert rel(fused, plain) < 1e-5
