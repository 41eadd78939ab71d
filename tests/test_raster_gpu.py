"""GPU parity of the rasteriser: bit-exact fragment indices / depth / payload against the C oracle
(oracle/rasterize_oracle.c) and the goldens produced by the UNMODIFIED reference kernels executed on the CPU
(tests/golden/raster_cases.npz), the reference's own golden mesh (body_vis.obj), and the backward against autograd
over the oracle's formulas."""
import numpy as np
import pytest
import torch

import golden_util as gu
import raster_cases
from oracle import rasterize_oracle as RO

pytestmark = pytest.mark.gpu


def run_cuda(fv, colors, h, w, dev, init=None):
    from gif_b200 import rasterize as R
    B = fv.shape[0]
    depth = torch.full((B, h, w), 1e6, device=dev) if init is None else torch.from_numpy(init[0]).to(dev)
    tri = torch.full((B, h, w), -1, dtype=torch.int32, device=dev) if init is None else torch.from_numpy(init[1]).to(dev)
    out3 = torch.zeros((B, h, w, 3), device=dev) if init is None else torch.from_numpy(init[2]).to(dev)
    fvt = torch.from_numpy(np.ascontiguousarray(fv)).to(dev)
    if colors is None:
        res = R.standard_rasterize(fvt, depth, tri, out3, h, w)
    else:
        res = R.standard_rasterize_colors(fvt, torch.from_numpy(colors).to(dev), depth, tri, out3, h, w)
    assert res[0] is depth and res[1] is tri and res[2] is out3          # in-place contract of the reference
    return depth.cpu().numpy(), tri.cpu().numpy(), out3.cpu().numpy()


@pytest.mark.parametrize("name", list(raster_cases.all_cases().keys()))
def test_raster_cases_bit_exact(cuda, name):
    fv, colors, h, w = raster_cases.all_cases()[name]
    g = gu.load_golden("raster_cases.npz")
    d, t, o = run_cuda(fv, colors, h, w, cuda)
    assert np.array_equal(d, g[name + "_depth"]), "depth differs from the reference kernels"
    assert np.array_equal(t, g[name + "_tri"]), "fragment index differs (lowest-index tie-break policy)"
    assert np.array_equal(o, g[name + "_out3"]), "barycentric / colour payload differs"
    # live oracle too
    d2, t2, o2 = (RO.oracle_rasterize_colors(fv, colors, h, w) if colors is not None else RO.oracle_rasterize(fv, h, w))
    assert np.array_equal(d, d2) and np.array_equal(t, t2) and np.array_equal(o, o2)


def test_raster_in_place_depth_test(cuda):
    """Rasterising into a pre-populated depth buffer: only nearer fragments replace (atomicMin semantics, :150-160)."""
    fv, _, h, w = raster_cases.all_cases()["soup_b2_64"]
    rng = np.random.default_rng(0)
    d0 = rng.uniform(1.0, 3.0, (2, h, w)).astype(np.float32)
    t0 = np.full((2, h, w), 7777, np.int32)
    o0 = rng.uniform(0, 1, (2, h, w, 3)).astype(np.float32)
    want = RO.oracle_rasterize(fv, h, w, d0.copy(), t0.copy(), o0.copy())
    got = run_cuda(fv, None, h, w, cuda, init=(d0.copy(), t0.copy(), o0.copy()))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert (got[1] == 7777).any() and (got[1] != 7777).any()


def test_body_visibility_golden(cuda):
    """The reference's checked-in golden: data/obj/body.obj * 0.8 at 512^2 -> per-vertex visibility of body_vis.obj."""
    from gif_b200 import rasterize as R
    g = gu.load_golden("body_visibility.npz")
    v = torch.from_numpy(g["vertices"])[None].to(cuda)
    f = torch.from_numpy(g["faces"])[None].to(cuda)
    vis = R.get_visibility(v, f, 512, 512)
    assert int((vis[0].cpu().numpy() != g["visible"]).sum()) == 0


def test_flame_batch_matches_oracle(cuda):
    """BASELINE configs[3] shape: FLAME topology (F=9976), 256^2, a batch of 8 random poses -- bit-exact vs the oracle,
    and the list-overflow fallback (tiny workspace is not reachable through the API, so exercise big triangles)."""
    from gif_b200.flame_synth import synthetic_flame_batch
    fv, colors = synthetic_flame_batch(8, 256, 256, seed=3, device="cpu")
    d, t, o = run_cuda(fv.numpy(), colors.numpy(), 256, 256, cuda)
    d2, t2, o2 = RO.oracle_rasterize_colors(fv.numpy(), colors.numpy(), 256, 256)
    assert np.array_equal(t, t2) and np.array_equal(d, d2) and np.array_equal(o, o2)
    assert (t >= 0).mean() > 0.2
    # overflow fallback: 3000 image-sized triangles per image overflow the per-image list capacity (8F+1024 slots;
    # each triangle touches up to 64 bins) -> brute-force path for that image, still bit-exact
    rng = np.random.default_rng(1)
    big = np.zeros((1, 3000, 3, 3), np.float32)
    big[..., :2] = rng.uniform(-20, 84, (1, 3000, 3, 2))
    big[..., 2] = rng.uniform(1, 2, (1, 3000, 3))
    d, t, o = run_cuda(big, None, 64, 64, cuda)
    d2, t2, o2 = RO.oracle_rasterize(big, 64, 64)
    assert np.array_equal(t, t2) and np.array_equal(d, d2) and np.array_equal(o, o2)


@pytest.mark.parametrize("with_colors", [False, True])
def test_raster_backward(cuda, with_colors):
    from gif_b200 import rasterize as R
    fv, colors = raster_cases.random_soup(2, 300, 48, 48, 11, size=7.0, with_colors=with_colors)
    h = w = 48
    fvt = torch.from_numpy(fv).to(cuda).requires_grad_(True)
    ct = torch.from_numpy(colors).to(cuda).requires_grad_(True) if with_colors else None
    depth, tri, out3 = R.rasterize(fvt, h, w, ct)
    rng = np.random.default_rng(5)
    g_out = torch.from_numpy(rng.normal(size=(2, h, w, 3)).astype(np.float32))
    g_dep = torch.from_numpy(rng.normal(size=(2, h, w)).astype(np.float32))
    mask = (tri >= 0).cpu()
    loss = (out3 * g_out.to(cuda)).sum() + (depth * (g_dep * mask).to(cuda)).sum()
    grads = torch.autograd.grad(loss, [fvt] + ([ct] if with_colors else []))
    # oracle: float64 autograd over the same formulas at the winners chosen by the forward
    fvo = torch.from_numpy(fv).double().requires_grad_(True)
    co = torch.from_numpy(colors).double().requires_grad_(True) if with_colors else None
    bw, dep, img = RO.interp_torch(fvo, tri.cpu(), co)
    payload = img if with_colors else bw
    loss_o = (payload * g_out.double()).sum() + (dep * (g_dep * mask).double()).sum()
    grads_o = torch.autograd.grad(loss_o, [fvo] + ([co] if with_colors else []))
    for a, b in zip(grads, grads_o):
        a, b = a.cpu().double().numpy(), b.numpy()
        # thin slivers have huge, ill-conditioned gradients: compare on the well-conditioned faces
        scale = np.abs(b).reshape(b.shape[0], b.shape[1], -1).max(-1)
        ok = scale < np.percentile(scale, 98)
        assert np.abs(a[ok] - b[ok]).max() / max(np.abs(b[ok]).max(), 1e-30) < 2e-3
