"""CPU: the oracle restatement (oracle/stylegan2_oracle.py) against the golden vectors produced by the unmodified
reference modules (oracle/make_golden.py).  Runs everywhere (no GPU, no /root/reference)."""

import torch

import golden_util as gu
from oracle import stylegan2_oracle as O

UPFIRDN_CASES = [(1, 1, (1, 1), 4.0), (1, 1, (2, 2), 1.0), (1, 1, (1, 1), 1.0), (2, 1, (2, 1), 4.0),
                 (1, 2, (1, 1), 1.0), (1, 2, (0, 0), 1.0), (1, 1, (-1, 0), 1.0)]


def test_upfirdn2d():
    g = gu.load_golden("ops.npz")
    for ci, (up, down, pad, gain) in enumerate(UPFIRDN_CASES):
        for si, (h, w) in enumerate([(9, 9), (8, 5), (17, 33)]):
            y = O.upfirdn2d(gu.randn((2, 3, h, w), 10 + ci * 7 + si), gu.blur_kernel(gain), up, down, pad)
            assert gu.rel_err(y.numpy(), g[f"upfirdn_{ci}_{si}"]) < 2e-6
    y = O.upfirdn2d(gu.randn((1, 2, 7, 6), 98), torch.from_numpy(g["upfirdn_asym_k"]), 2, 1, (2, 1))
    assert gu.rel_err(y.numpy(), g["upfirdn_asym"]) < 2e-6


def test_upfirdn2d_adjoint_rule():
    """The adjoint rule the CUDA backward uses (swap up/down, reversed kernel, pad0' = k-1-pad0, output size = forward
    input size) equals autograd of the oracle -- pins gif_b200.ops._UpFirDn.backward on the CPU."""
    k = gu.randn((4, 4), 5)
    for up, down, pad, hw in [(1, 1, (2, 2), (10, 7)), (2, 1, (2, 1), (6, 5)), (1, 2, (1, 1), (10, 8)), (1, 2, (0, 0), (9, 7))]:
        x = gu.randn((1, 2) + hw, 6).requires_grad_(True)
        y = O.upfirdn2d(x, k, up, down, pad)
        gy = gu.randn(tuple(y.shape), 7)
        (gx,) = torch.autograd.grad(y, x, gy)
        # adjoint as an upfirdn2d call producing exactly the input size: pad1' chosen to give that size
        kh = 4
        q0 = kh - 1 - pad[0]
        q1 = hw[0] * up - (gy.shape[2] * down + q0 - kh + 1) + (up - 1)   # solve out size == hw for axis 0 (then crop)
        adj = O.upfirdn2d(gy, torch.flip(k, [0, 1]), down, up, (q0, max(q1, 0) + 4))[:, :, :hw[0], :hw[1]]
        assert gu.rel_err(adj.numpy(), gx.numpy()) < 1e-5


def test_small_ops():
    g = gu.load_golden("ops.npz")
    y = O.fused_leaky_relu(gu.randn((2, 5, 4, 4), 4), gu.randn((1, 5, 1, 1), 3))
    assert gu.rel_err(y.numpy(), g["lrelu"]) < 1e-7
    y = O.equal_linear(gu.randn((3, 24), 7), gu.randn((16, 24), 5) * 100.0, gu.randn((16,), 6), 0.01, True)
    assert gu.rel_err(y.numpy(), g["linear_act"]) < 2e-6
    y = O.equal_linear(gu.randn((3, 24), 7), gu.randn((16, 24), 5), gu.randn((16,), 6))
    assert gu.rel_err(y.numpy(), g["linear_plain"]) < 2e-6
    for tag, (k, s, p, hw) in {"k1": (1, 1, 0, 8), "k3s1": (3, 1, 1, 8), "k3s2": (3, 2, 0, 9), "k1s2": (1, 2, 0, 7)}.items():
        y = O.equal_conv2d(gu.randn((2, 6, hw, hw), 9), gu.randn((10, 6, k, k), 8), None, s, p)
        assert gu.rel_err(y.numpy(), g[f"conv_{tag}"]) < 2e-6


def test_modulated_conv():
    g = gu.load_golden("modconv.npz")
    for tag, (ci, co, k, demod, up, hw) in {"plain": (32, 16, 3, True, False, 8), "up": (16, 32, 3, True, True, 5),
                                            "rgb": (32, 3, 1, False, False, 8)}.items():
        w = gu.randn((1, co, ci, k, k), 20).requires_grad_(True)
        mw = gu.randn((ci, 512), 21).requires_grad_(True)
        mb = (1.0 + 0.1 * gu.randn((ci,), 22)).requires_grad_(True)
        x = gu.randn((3, ci, hw, hw), 23).requires_grad_(True)
        st = gu.randn((3, 512), 24).requires_grad_(True)
        y = O.modulated_conv2d(x, st, w, mw, mb, demod, up, gu.blur_kernel(4.0) if up else None)
        assert gu.rel_err(y.detach().numpy(), g[f"{tag}_y"]) < 2e-5
        grads = torch.autograd.grad((y * gu.randn(tuple(y.shape), 25)).sum(), [x, st, w, mw, mb])
        for n, gr in zip("x style w modw modb".split(), grads):
            assert gu.rel_err(gr.numpy(), g[f"{tag}_g{n}"]) < 5e-5, (tag, n)


def test_generator_step3():
    g = gu.load_golden("generator.npz")
    sd = gu.seeded_state_dict(gu.g_shapes(100), 1)
    cond = gu.rand_uniform((2, 6, 32, 32), 40)
    with torch.no_grad():
        img = O.generator_forward(cond, gu.randint(100, (2,), 41), sd, step=3)
        assert gu.rel_err(img.numpy(), g["s3_img"]) < 5e-5
        img_z = O.generator_forward(cond, gu.randn((2, 512), 43), sd, step=3)
        assert gu.rel_err(img_z.numpy(), g["s3_img_z"]) < 5e-5


def test_discriminator_64_with_r1():
    g = gu.load_golden("discriminator.npz")
    sd = gu.seeded_state_dict(gu.d_shapes(64), 2)
    img = gu.rand_uniform((8, 3, 64, 64), 50).requires_grad_(True)
    cond = gu.rand_uniform((8, 6, 64, 64), 51)
    sc = O.discriminator_forward(img, cond, sd, 64)
    pen = O.r1_penalty(sc, img)
    assert gu.rel_err(sc.detach().numpy(), g["d64_scores"]) < 5e-5
    assert gu.rel_err(pen.detach().numpy(), g["d64_r1"]) < 2e-4
    # float64 evaluation reproduces the float64 reference to rounding
    sd64 = {k: v.double() for k, v in sd.items()}
    img64 = img.detach().double().requires_grad_(True)
    sc64 = O.discriminator_forward(img64, cond.double(), sd64, 64)
    assert gu.rel_err(sc64.detach().numpy(), g["d64_scores_f64"]) < 1e-11
    assert gu.rel_err(O.r1_penalty(sc64, img64).detach().numpy(), g["d64_r1_f64"]) < 1e-11


def test_cond_pyramid_equals_bilinear():
    """power-of-two bilinear (align_corners=False) reduction == mean of the central 2x2."""
    x = gu.randn((2, 6, 64, 64), 3)
    for size in (64, 32, 16, 8, 4):
        ref = torch.nn.functional.interpolate(x, size=(size, size), mode="bilinear", align_corners=False)
        assert gu.rel_err(O.cond_pyramid_level(x, size).numpy(), ref.numpy()) < 1e-6
