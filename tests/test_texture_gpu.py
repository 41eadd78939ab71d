"""GPU: gifb200_texture_steal_{fwd,bwd} and the FlameTextureSpace / InterpolatedTextureLoss modules against the reference's
golden and the oracle (fp32, bar 1e-5 on values of O(1); the visibility mask must match bit for bit away from
|normal_z| < 1e-6)."""
import numpy as np
import pytest
import torch

import golden_util as gu
from gif_b200.flame_synth import synthetic_flame_model, synthetic_texture_data
from oracle import flame_oracle as FO
from oracle import texture_oracle as TO

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _space(cuda):
    from gif_b200.flame import FLAME
    from gif_b200.texture_space import FlameTextureSpace
    fl = FLAME.from_arrays(synthetic_flame_model()).to(cuda)
    return FlameTextureSpace(synthetic_texture_data(), flame=fl), fl


def test_compute_texture_map_matches_reference_golden(cuda):
    from gif_b200.render import batch_orth_proj, vertex_normals
    g = gu.load_golden("texture_steal.npz")
    ts, fl = _space(cuda)
    verts, cam, src = (torch.from_numpy(g[k]).to(cuda) for k in ("verts", "cam", "src"))
    tv = batch_orth_proj(verts, cam)
    tv[:, :, 1:] = -tv[:, :, 1:]
    tex, mask = ts.compute_texture_map(src, verts, vertex_normals(tv, fl.faces_tensor), camera_params=cam)
    assert tuple(tex.shape) == (3, 3, 256, 256) and mask.dtype == torch.bool and tuple(mask.shape) == (3, 1, 256, 256)
    assert np.abs(tex.cpu().numpy()[:, :, ::3, ::3] - g["tex_sub"]).max() < TOL
    ref_mask = np.unpackbits(g["mask"])[:mask.numel()].reshape(mask.shape).astype(bool)
    assert (mask.cpu().numpy() != ref_mask).mean() < 1e-4      # normal_z within rounding of 0 may flip


def test_forward_from_flame_params_and_gradient(cuda):
    """FlameTextureSpace.forward: FLAME decode + normals + stealing, vs the oracle chain; d(loss)/d(image) vs autograd
    through the oracle's grid_sample; second derivative closure (the adjoint's backward is the forward)."""
    m, td = synthetic_flame_model(), synthetic_texture_data()
    ts, fl = _space(cuda)
    gen = torch.Generator().manual_seed(9)
    B = 3
    params = torch.cat([torch.randn(B, 100, generator=gen), torch.randn(B, 50, generator=gen),
                        (torch.rand(B, 6, generator=gen) * 2 - 1) * torch.tensor([0.2, 0.6, 0.1, 0.3, 0.02, 0.02]),
                        torch.rand(B, 1, generator=gen) * 3 + 6, (torch.rand(B, 2, generator=gen) * 2 - 1) * 0.03], 1)
    src = torch.rand(B, 3, 40, 56, generator=gen) * 2 - 1
    sg = src.to(cuda).requires_grad_(True)
    tex, mask = ts(sg, params.to(cuda))
    verts_o, _, _ = FO.flame_forward(m, params[:, :100], params[:, 100:150], params[:, 150:156])
    so = src.clone().requires_grad_(True)
    tex_o, mask_o = TO.texture_space_forward(so, verts_o, params[:, 156:159], m["faces"], td)
    # chained test: the GPU decoder's vertices differ from the oracle's by ~1e-6 m; x cam scale 9 x 28 px half-width x a
    # white-noise image (slope up to 2 per pixel) -> a few 1e-5 on the sampled values
    assert (tex.detach().cpu() - tex_o.detach()).abs().max() < 2e-4
    assert (mask.cpu() != mask_o).float().mean() < 1e-4
    w = torch.randn(tex_o.shape, generator=gen)
    (g_gpu,) = torch.autograd.grad((tex * w.to(cuda)).sum(), sg, create_graph=True)
    (g_o,) = torch.autograd.grad((tex_o * w).sum(), so)
    assert (g_gpu.detach().cpu() - g_o).abs().max() < 1e-4 * max(1.0, float(g_o.abs().max()))
    # closure: d/dw' <g(w'), v> = S(v)
    v = torch.randn(src.shape, generator=gen)
    wg = w.to(cuda).requires_grad_(True)
    (g2,) = torch.autograd.grad((ts(sg, params.to(cuda))[0] * wg).sum(), sg, create_graph=True)
    (gg,) = torch.autograd.grad((g2 * v.to(cuda)).sum(), wg)
    with torch.no_grad():
        sv, _ = TO.texture_space_forward(v, verts_o, params[:, 156:159], m["faces"], td)
    assert (gg.cpu() - sv).abs().max() < 2e-4


def test_interpolated_texture_loss_matches_oracle(cuda):
    """tex_sp_intrp_loss end to end on a small generator: same pairs, same identity -> same loss and same gradient
    w.r.t. a generator parameter as the oracle composition (oracle generator + oracle stealing + oracle loss)."""
    from gif_b200 import ops
    from gif_b200.model.stg2_generator import StyledGenerator
    from gif_b200.texture_space import InterpolatedTextureLoss
    from oracle import stylegan2_oracle as O
    ops.set_precision("fp32")
    try:
        m, td = synthetic_flame_model(), synthetic_texture_data()
        ts, fl = _space(cuda)
        sd = gu.seeded_state_dict(gu.g_shapes(16), 11)
        G = StyledGenerator(embedding_vocab_size=16, rendered_flame_ascondition=True, normal_maps_as_cond=True)
        G.load_state_dict(sd)
        G.to(cuda)
        gen = torch.Generator().manual_seed(4)
        n = 4
        flame_batch = torch.cat([torch.randn(n, 100, generator=gen), torch.randn(n, 50, generator=gen),
                                 (torch.rand(n, 6, generator=gen) * 2 - 1) * 0.3,
                                 torch.rand(n, 1, generator=gen) * 3 + 6, (torch.rand(n, 2, generator=gen) * 2 - 1) * 0.03], 1)
        cond = gu.rand_uniform((n, 6, 16, 16), 8)
        region = (torch.rand(1, 1, 256, 256, generator=gen) > 0.3).float()

        class Rng:                                   # fixed "random" choices, recorded for the oracle
            def choice(self, a, k, replace=False):
                return np.arange(k) % a
            def randint(self, lo, hi):
                return 5
        L = InterpolatedTextureLoss(n + 1, ts, lambda fb: cond[:fb.shape[0]].to(cuda), region.to(cuda), rng=Rng())
        key = "generator.progression.2.st_cv2.conv.weight"
        p = dict(G.named_parameters())[key]
        p.requires_grad_(True)
        loss = L.tex_sp_intrp_loss(flame_batch.to(cuda), lambda x, **kw: G(x, **kw), step=2, alpha=1, max_ids=16)
        (gp,) = torch.autograd.grad(loss, p)
        # oracle
        sdo = {k: v.clone() for k, v in sd.items()}
        sdo[key].requires_grad_(True)
        img_o = O.generator_forward(cond, torch.full((n,), 5, dtype=torch.long), sdo, step=2)
        verts_o, _, _ = FO.flame_forward(m, flame_batch[:, :100], flame_batch[:, 100:150], flame_batch[:, 150:156])
        tex_o, mask_o = TO.texture_space_forward(img_o, verts_o, flame_batch[:, 156:159], m["faces"], td)
        pairs = L.pairs[np.arange(L.max_num) % len(L.pairs)]
        loss_o = TO.tex_sp_intrp_loss(tex_o, mask_o.float(), [tuple(q) for q in pairs], region)
        (gpo,) = torch.autograd.grad(loss_o, sdo[key])
        assert abs(float(loss.detach()) - float(loss_o.detach())) < 1e-4 * abs(float(loss_o.detach()))
        assert gu.rel_err(gp.cpu().numpy(), gpo.numpy()) < 5e-3
    finally:
        ops.set_precision("tf32")
