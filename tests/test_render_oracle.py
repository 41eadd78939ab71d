"""CPU: oracle/render_oracle.py against goldens produced by the reference's own util.vertex_normals / batch_orth_proj and
Renderer.add_SHlight (oracle/make_raster_golden.py::render_pieces)."""
import numpy as np
import torch

import golden_util as gu
from oracle import render_oracle as RD


def test_render_pieces_match_reference_goldens():
    g = gu.load_golden("render_pieces.npz")
    faces = torch.from_numpy(gu.load_golden("flame_template.npz")["faces"].astype(np.int64))
    verts, cam = torch.from_numpy(g["verts"]), torch.from_numpy(g["cam"])
    assert np.abs(RD.vertex_normals(verts, faces).numpy() - g["normals"]).max() < 1e-6
    assert np.abs(RD.batch_orth_proj(verts, cam).numpy() - g["proj"]).max() < 1e-6
    s = RD.add_sh_light(torch.from_numpy(g["nimg"]), torch.from_numpy(g["sh"]))
    assert np.abs(s.numpy() - g["shading"]).max() < 1e-5


def test_composite_shade_matches_reference_renderer_golden():
    """oracle/render_oracle.shade (attribute interpolation, albedo grid_sample, SH shading, alpha, quantisation) against the
    UNMODIFIED reference Renderer.forward / render_normal driven as gif_helper does (tests/golden/render_composite.npz,
    oracle/make_render_golden.py).  Both sit on the same rasterisation (the pytorch3d-convention oracle), so everything the
    reference's own code computes downstream of (pix_to_face, bary) is pinned here."""
    from oracle import rasterize_oracle as RO
    g = gu.load_golden("render_composite.npz")
    z = gu.load_golden("flame_template.npz")
    faces = torch.from_numpy(z["faces"].astype(np.int64))
    uv, uvf = torch.from_numpy(z["uvcoords"]).float(), torch.from_numpy(z["uvfaces"].astype(np.int64))
    verts, cam = torch.from_numpy(g["verts"]), torch.from_numpy(g["cam"])
    alb, lights = torch.from_numpy(g["albedo"]), torch.from_numpy(g["lights"])
    S = g["images"].shape[-1]
    tv = RD.batch_orth_proj(verts, cam)
    tv[:, :, 1:] = -tv[:, :, 1:]
    tv[:, :, 2] += 10
    tv[..., :2] = -tv[..., :2]                                           # renderer.py:55
    _, t, b = RO.oracle_rasterize_pytorch3d(tv[:, faces].numpy(), S, S)
    b = b * (t >= 0)[..., None]
    uvg = torch.cat([uv, torch.ones_like(uv[:, :1])], -1) * 2 - 1         # renderer.py:105-107
    uvg[:, 1] = -uvg[:, 1]
    n = RD.vertex_normals(verts, faces)
    img, nrm, cond = RD.shade(torch.from_numpy(t), torch.from_numpy(b), uvg[uvf][:, :, :2], n[:, faces], alb, lights)
    assert np.array_equal((t >= 0)[:, None].astype(np.float32), g["alpha"])
    assert gu.rel_err(img.numpy(), g["images"]) < 2e-5
    assert gu.rel_err(nrm.numpy(), g["normal_images"]) < 2e-5
    tq = (cond[:, :3] + 1) / 2
    nq = (cond[:, 3:] + 1) / 2
    # floor() of a value that differs in the last float bit can move one quantisation level on isolated pixels
    assert (np.abs(tq.numpy() - np.clip(g["tex_quantised"], 0, 1)) > 1e-6).mean() < 1e-4
    assert (np.abs(nq.numpy() - g["normal_quantised"]) > 1e-6).mean() < 1e-4
