"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz`` by running the UNMODIFIED reference modules
(imported from /root/reference through ``oracle/ref_import.py``) on seeded inputs, and at the same time pins
the restatement in ``oracle/stylegan2_oracle.py`` against them (asserts, fp32 and fp64).

Run in the build container only:   python -m oracle.make_golden
The fixtures it writes are committed; the GPU box never needs /root/reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_util as gu  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle import stylegan2_oracle as O  # noqa: E402

torch.set_num_threads(8)
R = ref_import.load()
OUT = gu.GOLDEN_DIR
os.makedirs(OUT, exist_ok=True)
report = []


def check(name, a, b, tol):
    e = gu.rel_err(a.detach().double().numpy(), b.detach().double().numpy())
    report.append((name, e))
    assert e < tol, f"oracle != reference for {name}: rel err {e:.3e} (tol {tol})"
    return e


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **{k: np.asarray(v) for k, v in arrs.items()})


# ------------------------------------------------------------------ operators
def ops_golden():
    out = {}
    # upfirdn2d: every (up, down, pad, gain) instance on the hot path + ragged sizes
    cases = [(1, 1, (1, 1), 4.0), (1, 1, (2, 2), 1.0), (1, 1, (1, 1), 1.0), (2, 1, (2, 1), 4.0),
             (1, 2, (1, 1), 1.0), (1, 2, (0, 0), 1.0), (1, 1, (-1, 0), 1.0)]
    for ci, (up, down, pad, gain) in enumerate(cases):
        for si, (h, w) in enumerate([(9, 9), (8, 5), (17, 33)]):
            x = gu.randn((2, 3, h, w), 10 + ci * 7 + si)
            k = gu.blur_kernel(gain)
            y_ref = R.cl.upfirdn2d(x, k, up=up, down=down, pad=pad)
            y_o = O.upfirdn2d(x, k, up=up, down=down, pad=pad)
            check(f"upfirdn2d[{ci},{si}]", y_o, y_ref, 2e-6)
            out[f"upfirdn_{ci}_{si}"] = y_ref.numpy()
    # an asymmetric kernel pins the flip convention (cl.py:64)
    ka = gu.randn((4, 4), 99)
    x = gu.randn((1, 2, 7, 6), 98)
    y_ref = R.cl.upfirdn2d(x, ka, up=2, down=1, pad=(2, 1))
    check("upfirdn2d[asym]", O.upfirdn2d(x, ka, up=2, down=1, pad=(2, 1)), y_ref, 2e-6)
    out["upfirdn_asym"] = y_ref.numpy()
    out["upfirdn_asym_k"] = ka.numpy()

    # FusedLeakyReLU
    m = R.cl.FusedLeakyReLU(5)
    m.bias.data = gu.randn((1, 5, 1, 1), 3)
    x = gu.randn((2, 5, 4, 4), 4)
    y = m(x)
    check("fused_lrelu", O.fused_leaky_relu(x, m.bias.data), y, 1e-7)
    out["lrelu"] = y.detach().numpy()

    # EqualLinear (activated lr_mul 0.01 as in z_to_w; plain as in modulation)
    for tag, kw in (("act", dict(lr_mul=0.01, activation="fused_lrelu")), ("plain", dict(bias_init=1))):
        m = R.cl.EqualLinear(24, 16, **kw)
        m.weight.data = gu.randn((16, 24), 5) * (100.0 if tag == "act" else 1.0)
        m.bias.data = gu.randn((16,), 6)
        x = gu.randn((3, 24), 7)
        y = m(x)
        check(f"equal_linear[{tag}]", O.equal_linear(x, m.weight.data, m.bias.data, kw.get("lr_mul", 1.0),
                                                     tag == "act"), y, 2e-6)
        out[f"linear_{tag}"] = y.detach().numpy()

    # EqualConv2d
    for tag, (k, s, p, hw) in {"k1": (1, 1, 0, 8), "k3s1": (3, 1, 1, 8), "k3s2": (3, 2, 0, 9), "k1s2": (1, 2, 0, 7)}.items():
        m = R.cl.EqualConv2d(6, 10, k, stride=s, padding=p, bias=False)
        m.weight.data = gu.randn((10, 6, k, k), 8)
        x = gu.randn((2, 6, hw, hw), 9)
        y = m(x)
        check(f"equal_conv[{tag}]", O.equal_conv2d(x, m.weight.data, None, s, p), y, 2e-6)
        out[f"conv_{tag}"] = y.detach().numpy()
    save("ops.npz", **out)


def modconv_golden():
    """ModulatedConv2d fwd + all first-order grads: plain 3x3, upsample 3x3, 1x1 no-demod (ToRGB)."""
    out = {}
    for tag, (ci, co, k, demod, up, hw) in {"plain": (32, 16, 3, True, False, 8), "up": (16, 32, 3, True, True, 5),
                                            "rgb": (32, 3, 1, False, False, 8)}.items():
        m = R.cl.ModulatedConv2d(ci, co, k, 512, demodulate=demod, upsample=up)
        m.weight.data = gu.randn((1, co, ci, k, k), 20)
        m.modulation.weight.data = gu.randn((ci, 512), 21)
        m.modulation.bias.data = 1.0 + 0.1 * gu.randn((ci,), 22)
        x = gu.randn((3, ci, hw, hw), 23).requires_grad_(True)
        st = gu.randn((3, 512), 24).requires_grad_(True)
        y = m(x, st)
        gy = gu.randn(tuple(y.shape), 25)
        grads = torch.autograd.grad((y * gy).sum(), [x, st, m.weight, m.modulation.weight, m.modulation.bias])
        # oracle (fp32) vs reference
        w_o = m.weight.data.clone().requires_grad_(True)
        mw_o = m.modulation.weight.data.clone().requires_grad_(True)
        mb_o = m.modulation.bias.data.clone().requires_grad_(True)
        x_o = x.detach().clone().requires_grad_(True)
        st_o = st.detach().clone().requires_grad_(True)
        y_o = O.modulated_conv2d(x_o, st_o, w_o, mw_o, mb_o, demod, up, gu.blur_kernel(4.0) if up else None)
        g_o = torch.autograd.grad((y_o * gy).sum(), [x_o, st_o, w_o, mw_o, mb_o])
        check(f"modconv[{tag}].y", y_o, y, 2e-5)
        for n, a, b in zip("x style w modw modb".split(), g_o, grads):
            check(f"modconv[{tag}].g{n}", a, b, 5e-5)
        out[f"{tag}_y"] = y.detach().numpy()
        for n, g in zip("x style w modw modb".split(), grads):
            out[f"{tag}_g{n}"] = g.numpy()
    # BASELINE config 1: ModulatedConv2d(512,512,3,512), x (4,512,64,64), seed 0 -- sampled
    torch.manual_seed(0)
    m = R.cl.ModulatedConv2d(512, 512, 3, 512)
    m.weight.data = gu.randn((1, 512, 512, 3, 3), 30)
    m.modulation.weight.data = gu.randn((512, 512), 31)
    x = gu.randn((4, 512, 64, 64), 32)
    st = gu.randn((4, 512), 33)
    with torch.no_grad():
        y = m(x, st)
        y_o = O.modulated_conv2d(x, st, m.weight.data, m.modulation.weight.data, m.modulation.bias.data)
    check("modconv[config1].y", y_o, y, 2e-5)
    s, tot = gu.sample(y, 4096, 1)
    out["config1_sample"] = s
    out["config1_sum"] = tot
    out["config1_absmax"] = float(y.abs().max())
    save("modconv.npz", **out)


def generator_golden():
    """StyledGenerator forward (+ grads) at step 3 (32x32) full arrays, and step 6 (256x256) sampled."""
    shapes = gu.g_shapes(vocab=100)
    with ref_import.quiet():
        G = R.gen.StyledGenerator(embedding_vocab_size=100, rendered_flame_ascondition=True, normal_maps_as_cond=True,
                                  core_tensor_res=4, n_mlp=8)
    ref_shapes = {k: tuple(v.shape) for k, v in G.state_dict().items()}
    assert ref_shapes == {k: tuple(v) for k, v in shapes.items()}, "g_shapes manifest != reference state_dict"
    assert list(ref_shapes) == list(shapes), "g_shapes key order != reference state_dict order"
    sd = gu.seeded_state_dict(shapes, seed=1)
    G.load_state_dict(sd)
    out = {}
    # step 3, B=2, with backward
    cond = gu.rand_uniform((2, 6, 32, 32), 40).requires_grad_(True)
    idx = gu.randint(100, (2,), 41)
    img = G(cond, step=3, input_indices=idx)[0]
    gy = gu.randn(tuple(img.shape), 42)
    pnames = ["generator.progression.2.st_cv1.conv.weight", "generator.progression.3.st_cv2.conv.modulation.weight",
              "generator.progression.1.st_cv2.noise.noise_conv.4.weight", "generator.to_rgb.2.conv.weight",
              "generator.to_rgb.3.bias", "z_to_w.3.weight", "generator.const_input.input",
              "generator.progression.3.st_cv1.activate.bias"]
    named = dict(G.named_parameters())
    grads = torch.autograd.grad((img * gy).sum(), [cond] + [named[n] for n in pnames])
    sd_o = {k: v.clone().requires_grad_(k in pnames) for k, v in sd.items()}
    cond_o = cond.detach().clone().requires_grad_(True)
    img_o = O.generator_forward(cond_o, idx, sd_o, step=3)
    g_o = torch.autograd.grad((img_o * gy).sum(), [cond_o] + [sd_o[n] for n in pnames])
    check("G[step3].img", img_o, img, 5e-5)
    for n, a, b in zip(["cond"] + pnames, g_o, grads):
        check(f"G[step3].g[{n}] (fp32, mask-flip noise)", a, b, 3e-2)
    G64 = G.double()
    cond_d = cond.detach().double().requires_grad_(True)
    named_d = dict(G64.named_parameters())
    img_d = G64(cond_d, step=3, input_indices=idx)[0]
    g_d = torch.autograd.grad((img_d * gy.double()).sum(), [cond_d] + [named_d[n] for n in pnames])
    sd_q = {k: v.double().requires_grad_(k in pnames) for k, v in sd.items()}
    cond_q = cond.detach().double().requires_grad_(True)
    img_q = O.generator_forward(cond_q, idx, sd_q, step=3)
    g_q = torch.autograd.grad((img_q * gy.double()).sum(), [cond_q] + [sd_q[n] for n in pnames])
    check("G[step3].img fp64", img_q, img_d, 1e-11)
    for n, a, b in zip(["cond"] + pnames, g_q, g_d):
        check(f"G[step3].g[{n}] fp64", a, b, 1e-10)
    out["s3_img_f64"] = img_d.detach().numpy()
    out["s3_gcond_f64"] = g_d[0].numpy()
    for n, g in zip(pnames, g_d[1:]):
        s, tot = gu.sample(g, 2048, 2)
        out["s3_g64_" + n] = s
        out["s3_g64sum_" + n] = tot
    G.float()
    out["s3_img"] = img.detach().numpy()
    out["s3_gcond"] = grads[0].numpy()
    for n, g in zip(pnames, grads[1:]):
        s, tot = gu.sample(g, 2048, 2)
        out["s3_g_" + n] = s
        out["s3_gsum_" + n] = tot
    # float z path (gen.py:272-273)
    z = gu.randn((2, 512), 43)
    with torch.no_grad():
        img_z = G(cond.detach(), step=3, input_indices=z)[0]
        check("G[step3,z].img", O.generator_forward(cond.detach(), z, sd, step=3), img_z, 5e-5)
    out["s3_img_z"] = img_z.numpy()
    # step 6 (256^2), B=2, forward only, sampled
    cond = gu.rand_uniform((2, 6, 256, 256), 44)
    idx = gu.randint(100, (2,), 45)
    with torch.no_grad():
        img = G(cond, step=6, input_indices=idx)[0]
        img_o = O.generator_forward(cond, idx, sd, step=6)
    check("G[step6].img", img_o, img, 1e-4)
    s, tot = gu.sample(img, 8192, 3)
    out["s6_sample"] = s
    out["s6_sum"] = tot
    out["s6_absmax"] = float(img.abs().max())
    save("generator.npz", **out)


def discriminator_golden():
    """Discriminator(64, 9ch) B=8 fwd + bwd + R1 (double backward), full arrays; D(256) B=4 fwd sampled."""
    out = {}
    shapes = gu.d_shapes(64)
    with ref_import.quiet():
        D = R.disc.Discriminator(64, num_color_chnls=9)
    ref_shapes = {k: tuple(v.shape) for k, v in D.state_dict().items()}
    assert ref_shapes == {k: tuple(v) for k, v in shapes.items()}, "d_shapes manifest != reference state_dict"
    assert list(ref_shapes) == list(shapes)
    sd = gu.seeded_state_dict(shapes, seed=2)
    D.load_state_dict(sd)
    img = gu.rand_uniform((8, 3, 64, 64), 50).requires_grad_(True)
    cond = gu.rand_uniform((8, 6, 64, 64), 51).requires_grad_(True)
    pnames = ["convs.0.0.weight", "convs.1.conv1.0.weight", "convs.2.conv2.1.weight", "convs.3.skip.1.weight",
              "convs.2.conv2.2.bias", "final_conv.0.weight", "final_linear.0.weight", "final_linear.1.bias"]
    named = dict(D.named_parameters())
    scores, _ = D([img], condition=cond)
    pen = R.losses.grad_penalty_loss([img], scores, step=None)
    loss = torch.nn.functional.softplus(-scores).mean() + pen.mean()
    grads = torch.autograd.grad(loss, [img, cond] + [named[n] for n in pnames])

    sd_o = {k: v.clone().requires_grad_(k in pnames) for k, v in sd.items()}
    img_o = img.detach().clone().requires_grad_(True)
    cond_o = cond.detach().clone().requires_grad_(True)
    sc_o = O.discriminator_forward(img_o, cond_o, sd_o, 64)
    pen_o = O.r1_penalty(sc_o, img_o)
    loss_o = torch.nn.functional.softplus(-sc_o).mean() + pen_o.mean()
    g_o = torch.autograd.grad(loss_o, [img_o, cond_o] + [sd_o[n] for n in pnames])
    check("D[64].scores", sc_o, scores, 5e-5)
    check("D[64].r1", pen_o, pen, 2e-4)
    # fp32 gradients through ~14 leaky-ReLUs are only piecewise continuous: a pre-activation within rounding
    # noise of 0 flips its mask between two *correct* fp32 evaluation orders, so fp32-vs-fp32 agreement is loose
    # (the reference disagrees with its own fp64 evaluation by ~7e-3 max-norm).  fp64 arbitrates below.
    for n, a, b in zip(["img", "cond"] + pnames, g_o, grads):
        check(f"D[64].g[{n}] (fp32, mask-flip noise)", a, b, 3e-2)
    D64 = D.double()
    img_d = img.detach().double().requires_grad_(True)
    cond_d = cond.detach().double().requires_grad_(True)
    named_d = dict(D64.named_parameters())
    sc_d, _ = D64([img_d], condition=cond_d)
    pen_d = R.losses.grad_penalty_loss([img_d], sc_d, step=None)
    loss_d = torch.nn.functional.softplus(-sc_d).mean() + pen_d.mean()
    g_d = torch.autograd.grad(loss_d, [img_d, cond_d] + [named_d[n] for n in pnames])
    sd_q = {k: v.double().requires_grad_(k in pnames) for k, v in sd.items()}
    img_q = img.detach().double().requires_grad_(True)
    cond_q = cond.detach().double().requires_grad_(True)
    sc_q = O.discriminator_forward(img_q, cond_q, sd_q, 64)
    pen_q = O.r1_penalty(sc_q, img_q)
    loss_q = torch.nn.functional.softplus(-sc_q).mean() + pen_q.mean()
    g_q = torch.autograd.grad(loss_q, [img_q, cond_q] + [sd_q[n] for n in pnames])
    check("D[64].scores fp64", sc_q, sc_d, 1e-11)
    check("D[64].r1 fp64", pen_q, pen_d, 1e-11)
    for n, a, b in zip(["img", "cond"] + pnames, g_q, g_d):
        check(f"D[64].g[{n}] fp64", a, b, 1e-10)
    out["d64_scores_f64"] = sc_d.detach().numpy()
    out["d64_r1_f64"] = pen_d.detach().numpy()
    out["d64_gimg_f64"] = g_d[0].numpy()
    out["d64_gcond_f64"] = g_d[1].numpy()
    out["d64_gimg_ref32_l2err"] = float((grads[0].double() - g_d[0]).norm() / g_d[0].norm())
    for n, g in zip(pnames, g_d[2:]):
        s, tot = gu.sample(g, 2048, 4)
        out["d64_g64_" + n] = s
        out["d64_g64sum_" + n] = tot
    D.float()
    out["d64_scores"] = scores.detach().numpy()
    out["d64_r1"] = pen.detach().numpy()
    out["d64_gimg"] = grads[0].numpy()
    out["d64_gcond"] = grads[1].numpy()
    for n, g in zip(pnames, grads[2:]):
        s, tot = gu.sample(g, 2048, 4)
        out["d64_g_" + n] = s
        out["d64_gsum_" + n] = tot
    # 256^2
    shapes = gu.d_shapes(256)
    with ref_import.quiet():
        D = R.disc.Discriminator(256, num_color_chnls=9)
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    sd = gu.seeded_state_dict(shapes, seed=3)
    D.load_state_dict(sd)
    img = gu.rand_uniform((4, 3, 256, 256), 52)
    cond = gu.rand_uniform((4, 6, 256, 256), 53)
    with torch.no_grad():
        scores, _ = D([img], condition=cond)
        check("D[256].scores", O.discriminator_forward(img, cond, sd, 256), scores, 1e-4)
    out["d256_scores"] = scores.numpy()
    save("discriminator.npz", **out)


def fp64_check():
    """The restatement vs the reference in float64 (arbitrates fp32 disagreements): must agree to ~1e-12."""
    m = R.cl.ModulatedConv2d(8, 6, 3, 512, upsample=True).double()   # make_kernel is hard-wired fp32 (cl.py:84)
    x = torch.randn(2, 8, 5, 5, dtype=torch.float64)
    st = torch.randn(2, 512, dtype=torch.float64)
    y = m(x, st)
    y_o = O.modulated_conv2d(x, st, m.weight.data, m.modulation.weight.data, m.modulation.bias.data, True, True,
                             gu.blur_kernel(4.0).double())
    check("fp64 modconv[up]", y_o, y, 1e-12)


if __name__ == "__main__":
    ops_golden()
    modconv_golden()
    fp64_check()
    generator_golden()
    discriminator_golden()
    w = max(len(n) for n, _ in report)
    with open(os.path.join(OUT, "ORACLE_VS_REFERENCE.txt"), "w") as f:
        f.write("# oracle/stylegan2_oracle.py vs unmodified reference modules (norm-wise rel err), written by "
                "oracle/make_golden.py\n")
        for n, e in report:
            f.write(f"{n:<{w}}  {e:.3e}\n")
            print(f"{n:<{w}}  {e:.3e}")
    print("golden fixtures written to", OUT)
