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
