"""Synthetic FLAME-shaped workload for the rasteriser (BASELINE.json configs[3]).

The FLAME model itself (generic_model.pkl, texture space) is licence-gated and absent from the reference tree, so
"random FLAME params" drive a FLAME-*shaped* linear decoder on the real FLAME topology (V=5023, F=9976; template from
my_utils/photometric_optimization/data/head_template_mesh.obj, stored as tests/golden/flame_template.npz):
    verts = T + sum_j beta_j S_j,  beta ~ N(0,1) (B,150)
with a SMOOTH basis S_j(v) = a * d_j * cos(2 pi f_j . T_v + phi_j) (|f_j| <= 8 cycles/m, a = 0.5 mm, i.e. ~4 mm rms total
displacement, comparable to FLAME's shape+expression range).  (SURVEY 8d suggested i.i.d. per-vertex noise N(0,(2 mm)^2)
per coefficient; that sums to 24 mm rms of *uncorrelated* vertex noise, which crumples the mesh into image-sized slivers
(mean bbox 1800 px^2 at 256^2) and no longer resembles a FLAME render, so it is not used.)  It is followed by the reference's camera path: a random head rotation (cf. plots/generate_random_samples.py:107-108),
util.batch_orth_proj (util.py:73-83) + the y/z flip of gif_helper.py:26-27, and the pixel mapping of
visibility.py:38-40 (x*w/2+w/2, y*h/2+h/2, z-min(z)+1).
"""
import math
import os

import numpy as np
import torch

_TEMPLATE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                         "flame_template.npz")
_cache = {}


def flame_topology():
    if "t" not in _cache:
        z = np.load(_TEMPLATE)
        v = torch.from_numpy(z["vertices"].astype(np.float32))
        _cache["t"] = (v - v.mean(0, keepdim=True), torch.from_numpy(z["faces"].astype(np.int64)))
    return _cache["t"]


def flame_uv():
    """(uvcoords (5118,2), uvfaces (9976,3)) of the FLAME template."""
    z = np.load(_TEMPLATE)
    return torch.from_numpy(z["uvcoords"].astype(np.float32)), torch.from_numpy(z["uvfaces"].astype(np.int64))


def synthetic_flame_params(batch, seed=0):
    """World-space vertices (B,V,3), weak-perspective cameras (B,3), albedo textures (B,3,256,256) in 0..255 and SH lights
    (B,9,3) for the render benchmark: the smooth FLAME-shaped decoder below + random pose / camera / smooth random albedo."""
    tmpl, faces = flame_topology()
    g = torch.Generator().manual_seed(4321 + seed)
    verts = _decode(batch, g, tmpl)
    cam = torch.cat([torch.rand(batch, 1, generator=g) * 3 + 7, (torch.rand(batch, 2, generator=g) * 2 - 1) * 0.02], 1)
    low = torch.rand(batch, 3, 8, 8, generator=g) * 255
    albedo = torch.nn.functional.interpolate(low, size=(256, 256), mode="bilinear", align_corners=False)
    lights = torch.zeros(batch, 9, 3)
    lights[:, 0] = 3.0 + 0.3 * torch.randn(batch, 3, generator=g)
    lights[:, 1:] = 0.3 * torch.randn(batch, 8, 3, generator=g)
    return verts, cam, albedo, lights


def _decode(batch, g, tmpl):
    gb = torch.Generator().manual_seed(7)
    freq = (torch.rand(150, 3, generator=gb) * 2 - 1) * 8.0
    phase = torch.rand(150, 1, generator=gb) * (2 * math.pi)
    dirs = torch.nn.functional.normalize(torch.randn(150, 3, generator=gb), dim=1)
    wave = torch.cos(2 * math.pi * (freq @ tmpl.t()) + phase)
    basis = (0.0005 * wave[:, :, None] * dirs[:, None, :]).reshape(150, -1)
    beta = torch.randn(batch, 150, generator=g)
    verts = tmpl[None] + (beta @ basis).reshape(batch, -1, 3)
    yaw = (torch.rand(batch, generator=g) * 2 - 1) * (math.pi / 8)
    pitch = torch.rand(batch, generator=g) * (math.pi / 12)
    cy, sy, cp, sp = torch.cos(yaw), torch.sin(yaw), torch.cos(pitch), torch.sin(pitch)
    zero, one = torch.zeros_like(cy), torch.ones_like(cy)
    ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], -1).reshape(batch, 3, 3)
    rx = torch.stack([one, zero, zero, zero, cp, -sp, zero, sp, cp], -1).reshape(batch, 3, 3)
    return verts @ (rx @ ry).transpose(1, 2)


def synthetic_flame_batch(batch, h, w, seed=0, device="cuda"):
    """-> face_vertices (B,F,3,3) fp32 in pixel space, face_colors (B,F,3,3) in [0,1]."""
    tmpl, faces = flame_topology()
    g = torch.Generator().manual_seed(1234 + seed)
    verts = _decode(batch, g, tmpl)
    cam = torch.cat([torch.rand(batch, 1, generator=g) * 3 + 7, (torch.rand(batch, 2, generator=g) * 2 - 1) * 0.02], 1)
    proj = torch.cat([verts[..., :2] + cam[:, None, 1:], verts[..., 2:]], -1) * cam[:, None, 0:1]   # batch_orth_proj
    proj[..., 1:] = -proj[..., 1:]                                                                   # gif_helper.py:27
    pix = proj.clone()
    pix[..., 0] = proj[..., 0] * w / 2 + w / 2
    pix[..., 1] = proj[..., 1] * h / 2 + h / 2
    pix[..., 2] = proj[..., 2] - proj[..., 2].min() + 1
    vcol = torch.rand(batch, tmpl.shape[0], 3, generator=g)
    fv = pix[:, faces]          # (B,F,3,3)
    fc = vcol[:, faces]
    return fv.contiguous().to(device), fc.contiguous().to(device)
