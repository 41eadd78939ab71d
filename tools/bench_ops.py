#!/usr/bin/env python
"""Per-operator measurements for the BASELINE.json configs other than the headline step (profiles/ evidence):
  config 0  ModulatedConv2d(512,512,3) fwd, x (4,512,64,64)                      -> ms, TFLOP/s
  config 1  StyledGenerator fwd 256^2 bs16                                       -> images/s
  config 3  FLAME-topology rasterise texture+normal 256^2 bs64 fwd+bwd           -> renders/s, GB/s vs algorithmic bytes
  3b        north-star layer ModulatedConv2d(128,128,3) x (32,128,256,256) fwd/bwd -> TFLOP/s vs tf32 peak
  memory-bound kernels (upfirdn2d blur, fused bias/act, ToRGB)                   -> GB/s vs measured HBM peak
CUDA-event timing on the current stream, >= 3 warm-ups, inputs larger than L2 or L2 flushed between iterations.
Prints one JSON object per line."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

from gif_b200 import ops, rasterize  # noqa: E402
from gif_b200.flame_synth import synthetic_flame_batch  # noqa: E402
from gif_b200.model import stylegan2_common_layers as cl  # noqa: E402
from gif_b200.model.stg2_generator import StyledGenerator  # noqa: E402

dev = torch.device("cuda:0")
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    pass
HBM = PEAKS.get("hbm_gbs", 6650.0)
TF32_BURST = PEAKS.get("bf16_tflops", 1590.0) / 2
_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3, flush=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def out(**kw):
    print(json.dumps(kw), flush=True)


def modconv(ci, co, b, r, tag):
    m = cl.ModulatedConv2d(ci, co, 3, 512).to(dev)
    x = torch.randn(b, r, r, ci, device=dev).permute(0, 3, 1, 2)          # NCHW view of channels-last storage
    st = torch.randn(b, 512, device=dev)
    flops = 2.0 * b * r * r * ci * co * 9
    with torch.no_grad():
        ms = timeit(lambda: m(x, st), flush=True)
    xg = x.detach().requires_grad_(True)

    def fb():
        y = m(xg, st)
        y.backward(torch.ones_like(y))
    ms_fb = timeit(fb, iters=5, flush=True)
    # the contraction alone (input already modulated + rounded): gifb200_conv2d through the C ABI
    xs = ops._round_tf32_raw(torch.randn(b, r, r, ci, device=dev))
    w = torch.randn(9, co, ci, device=dev) / math.sqrt(9 * ci)
    ms_k = timeit(lambda: ops._conv_raw(xs, w, 3, ops.S1, False, False, (r, r)), flush=True)
    gy = ops._round_tf32_raw(torch.randn(b, r, r, co, device=dev))
    ms_w = timeit(lambda: ops._wgrad_raw(xs, gy, 3, ops.S1, False, False), flush=True)
    out(bench=tag, shape=[b, ci, r, r, co], module_fwd_ms=ms, module_fwd_tflops=flops / ms / 1e9,
        module_fwd_bwd_ms=ms_fb, module_fwd_bwd_tflops=3 * flops / ms_fb / 1e9,
        conv_kernel_ms=ms_k, conv_kernel_tflops=flops / ms_k / 1e9, conv_kernel_frac_of_tf32_burst=flops / ms_k / 1e9 / TF32_BURST,
        wgrad_kernel_ms=ms_w, wgrad_kernel_tflops=flops / ms_w / 1e9, tf32_peak_burst=TF32_BURST, l2="flushed")


def generator_fwd():
    G = StyledGenerator(embedding_vocab_size=70000, rendered_flame_ascondition=True, normal_maps_as_cond=True).to(dev)
    cond = torch.rand(16, 6, 256, 256, device=dev) * 2 - 1
    idx = torch.randint(0, 70000, (16,), device=dev)
    with torch.no_grad():
        ms = timeit(lambda: G(cond, step=6, input_indices=idx), iters=5)
    out(bench="config1_generator_fwd_256_bs16", ms=ms, images_per_s=16 / ms * 1e3, tflops=16 * 105.3 / ms)


def raster():
    """BASELINE configs[3]: FLAME-topology texture + normal render 256^2 bs64, forward and forward+backward.  Measured two
    ways: (a) ONE rasterisation interpolating both attribute sets (gifb200_rasterize_fwd_ex with face_colors2: the path the
    renderer uses), (b) two standard_rasterize_colors calls as a user of the reference's pybind API would issue them."""
    b, h, w = 64, 256, 256
    fv, fc = synthetic_flame_batch(b, h, w, seed=0, device=dev)
    F = fv.shape[1]
    fvg = fv.clone().requires_grad_(True)
    fcg = fc.clone().requires_grad_(True)
    normals = torch.rand_like(fc).requires_grad_(True)
    g1 = torch.randn(b, h, w, 3, device=dev)

    def fwd_one():
        return rasterize.rasterize(fv, h, w, fc, normals.detach())

    def fwd_bwd_one():
        d, t, im, nm = rasterize.rasterize(fvg, h, w, fcg, normals)
        ((im * g1).sum() + (nm * g1).sum()).backward()

    def fwd_two():
        return rasterize.rasterize(fv, h, w, fc), rasterize.rasterize(fv, h, w, normals.detach())

    def fwd_bwd_two():
        d, t, im = rasterize.rasterize(fvg, h, w, fcg)
        d2, t2, nm = rasterize.rasterize(fvg, h, w, normals)
        ((im * g1).sum() + (nm * g1).sum()).backward()
    ms_f, ms_fb = timeit(fwd_one, flush=True), timeit(fwd_bwd_one, iters=5, flush=True)
    ms_f2, ms_fb2 = timeit(fwd_two, flush=True), timeit(fwd_bwd_two, iters=5, flush=True)
    # algorithmic bytes per image (texture+normal): faces + colours + normals in, depth/tri + 2 images out;
    # backward: 2 image gradients + tri in, faces/colours/normals in, 3 gradients out
    fwd_bytes = 3 * F * 36 + h * w * (4 + 4 + 12 + 12)
    bwd_bytes = h * w * (12 + 12 + 4) + 3 * F * 36 + 3 * F * 36
    out(bench="config3_rasterize_flame_256_bs64", fwd_ms=ms_f, fwd_bwd_ms=ms_fb, renders_per_s_fwd=b / ms_f * 1e3,
        renders_per_s_fwd_bwd=b / ms_fb * 1e3, fwd_gbs=b * fwd_bytes / ms_f / 1e6, fwd_frac_of_hbm=b * fwd_bytes / ms_f / 1e6 / HBM,
        fwd_bwd_gbs=b * (fwd_bytes + bwd_bytes) / ms_fb / 1e6, fwd_bwd_frac_of_hbm=b * (fwd_bytes + bwd_bytes) / ms_fb / 1e6 / HBM,
        algorithmic_mb_per_image_fwd=fwd_bytes / 1e6, algorithmic_mb_per_image_bwd=bwd_bytes / 1e6, hbm_peak_gbs=HBM,
        two_call_form={"fwd_ms": ms_f2, "fwd_bwd_ms": ms_fb2},
        covered_fraction=float((rasterize.rasterize(fv, h, w, fc)[1] >= 0).float().mean()),
        note="one rasterisation, two attribute sets (colours + normals); includes the Python wrapper's buffer initialisation")


def render():
    """FLAME params -> 6-channel condition map (rasterise + fused shading), bs64 @256^2."""
    from gif_b200.flame_synth import flame_topology, flame_uv, synthetic_flame_params
    from gif_b200.render import FlameRenderer
    b, S = 64, 256
    verts, cam, alb, lights = (t.to(dev) for t in synthetic_flame_params(b, seed=0))
    _, faces = flame_topology()
    uv, uvf = flame_uv()
    R = FlameRenderer(faces, uv, uvf, image_size=S).to(dev)
    ms = timeit(lambda: R.render_tex_and_normal(verts, cam, alb, lights), flush=True)
    out(bench="config3_flame_condition_render_256_bs64", ms=ms, renders_per_s=b / ms * 1e3,
        note="projection + vertex normals (torch glue on 5023 vertices) + rasterise + fused shade -> tex, normal, cond maps")


def flame_pipeline():
    """FLAME parameters -> LBS (gifb200_flame_lbs) -> projection -> rasterise -> fused shade -> cond map, bs64."""
    from gif_b200 import flame as gflame
    from gif_b200.flame_synth import flame_uv, synthetic_flame_model, synthetic_flame_params
    from gif_b200.render import FlameRenderer
    b, S = 64, 256
    fl = gflame.FLAME.from_arrays(synthetic_flame_model()).to(dev)
    g = torch.Generator().manual_seed(0)
    shape, exp = torch.randn(b, 100, generator=g).to(dev), torch.randn(b, 50, generator=g).to(dev)
    pose = ((torch.rand(b, 6, generator=g) * 2 - 1) * torch.tensor([0.2, 0.5, 0.1, 0.3, 0.0, 0.0])).to(dev)
    _, cam, alb, lights = (t.to(dev) for t in synthetic_flame_params(b, seed=0))
    uv, uvf = flame_uv()
    R = FlameRenderer(fl.faces_tensor.cpu(), uv, uvf, image_size=S).to(dev)
    betas = torch.cat([shape, exp], 1)
    full_pose = torch.cat([pose[:, :3], torch.zeros(b, 3, device=dev), pose[:, 3:], torch.zeros(b, 6, device=dev)], 1)
    ms_lbs = timeit(lambda: gflame.lbs(betas, full_pose, fl._model()), flush=True)
    ms_fwd = timeit(lambda: fl(shape, exp, pose), flush=True)
    ms_all = timeit(lambda: R.render_tex_and_normal(fl.decode_vertices(shape, exp, pose)[0], cam, alb, lights), flush=True)
    V, NB, P = 5023, 150, 36
    basis_bytes = (NB + P) * 3 * V * 4 * (b // 8) + b * V * 12          # bases re-read once per 8-sample group (L2 hits)
    out(bench="flame_lbs_bs64", lbs_kernels_ms=ms_lbs, lbs_decodes_per_s=b / ms_lbs * 1e3, lbs_gbs_incl_l2_rereads=basis_bytes / ms_lbs / 1e6,
        flame_forward_with_landmarks_ms=ms_fwd, params_to_condition_map_ms=ms_all, condition_maps_per_s=b / ms_all * 1e3,
        note="FLAME params -> vertices (2 kernels) -> landmarks (torch glue) ; full: + projection, normals, rasterise, shade")


def texture_steal():
    """FlameTextureSpace.forward on 31 generated 256^2 images (the InterpolatedTextureLoss batch), fwd and fwd+bwd."""
    from gif_b200 import flame as gflame
    from gif_b200.flame_synth import synthetic_flame_model, synthetic_texture_data
    from gif_b200.texture_space import FlameTextureSpace
    b = 31
    fl = gflame.FLAME.from_arrays(synthetic_flame_model()).to(dev)
    ts = FlameTextureSpace(synthetic_texture_data(), flame=fl)
    g = torch.Generator().manual_seed(0)
    params = torch.cat([torch.randn(b, 100, generator=g), torch.randn(b, 50, generator=g), (torch.rand(b, 6, generator=g) * 2 - 1) * 0.3,
                        torch.rand(b, 1, generator=g) * 3 + 6, (torch.rand(b, 2, generator=g) * 2 - 1) * 0.03], 1).to(dev)
    img = torch.randn(b, 256, 256, 3, device=dev).permute(0, 3, 1, 2).requires_grad_(True)
    ms_f = timeit(lambda: ts(img.detach(), params), flush=True)

    def fb():
        tex, _ = ts(img, params)
        tex.sum().backward()
    ms_fb = timeit(fb, iters=5, flush=True)
    nbytes = b * (256 * 256 * 3 * 4 * 2 + 256 * 256)          # image in, texture out, mask out
    out(bench="texture_steal_31x256x256", fwd_ms=ms_f, fwd_bwd_ms=ms_fb, textures_per_s=b / ms_f * 1e3,
        fwd_gbs=nbytes / ms_f / 1e6, note="FLAME decode + projection + vertex normals (torch glue) + gifb200_texture_steal_fwd")


def memory_bound():
    x = torch.randn(32, 256, 256, 128, device=dev)
    k = torch.tensor([1., 3., 3., 1.], device=dev)
    k = torch.outer(k, k) / 64
    nbytes = x.numel() * 4
    ms = timeit(lambda: ops.upfirdn2d(x, k, 1, 1, (2, 2)))
    ob = 32 * 257 * 257 * 128 * 4
    out(bench="upfirdn2d_blur_pad22_32x256x256x128", ms=ms, gbs=(nbytes + ob) / ms / 1e6, frac_of_hbm=(nbytes + ob) / ms / 1e6 / HBM)
    ms = timeit(lambda: ops.upfirdn2d(x, k, 1, 2, (1, 1)))
    ob = 32 * 128 * 128 * 128 * 4
    out(bench="upfirdn2d_down2_pad11_32x256x256x128", ms=ms, gbs=(nbytes + ob) / ms / 1e6, frac_of_hbm=(nbytes + ob) / ms / 1e6 / HBM)
    bias = torch.randn(128, device=dev)
    d = torch.rand(32, 128, device=dev)
    add = torch.randn_like(x)
    ms = timeit(lambda: ops.bias_act(x, bias, 0.2, math.sqrt(2), rowscale=d, add=add))
    out(bench="bias_act_fused_tail_32x256x256x128", ms=ms, gbs=3 * nbytes / ms / 1e6, frac_of_hbm=3 * nbytes / ms / 1e6 / HBM)
    ms = timeit(lambda: ops.bias_act(x, bias, 0.2, math.sqrt(2)))
    out(bench="bias_act_plain_32x256x256x128", ms=ms, gbs=2 * nbytes / ms / 1e6, frac_of_hbm=2 * nbytes / ms / 1e6 / HBM)
    ws = torch.randn(32, 3, 128, device=dev)
    ms = timeit(lambda: ops.torgb(x, ws))
    out(bench="torgb_32x256x256x128", ms=ms, gbs=(nbytes + 32 * 65536 * 12) / ms / 1e6, frac_of_hbm=(nbytes + 32 * 65536 * 12) / ms / 1e6 / HBM)
    s = torch.rand(32, 128, device=dev)
    ms = timeit(lambda: ops.chan_scale(x, s, True))
    out(bench="chan_scale_32x256x256x128", ms=ms, gbs=2 * nbytes / ms / 1e6, frac_of_hbm=2 * nbytes / ms / 1e6 / HBM)


if __name__ == "__main__":
    ops.set_precision("tf32")
    which = sys.argv[1:] or ["northstar", "config0", "generator", "raster", "memory"]
    if "northstar" in which:
        modconv(128, 128, 32, 256, "northstar_modconv_128_128_256x256_bs32")
    if "config0" in which:
        modconv(512, 512, 4, 64, "config0_modconv_512_512_64x64_bs4")
    if "generator" in which:
        generator_fwd()
    if "raster" in which:
        raster()
        render()
        flame_pipeline()
        texture_steal()
    if "memory" in which:
        memory_bound()
