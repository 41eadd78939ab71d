"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, fp32/fp64) of the reference's shading path downstream of the
rasteriser: attribute interpolation (photometric_optimization/renderer.py:69-84), albedo grid_sample + SH shading +
composition (renderer.py:152-221), render_normal (:291-305), OverLayViz quantisation (visualize_flame_overlay.py:29-31) and
the consumer's [-1,1] mapping (losses.py:213-214); plus util.vertex_normals / batch_orth_proj (util.py:73-83,156-189).

Pinning: ``vertex_normals``, ``batch_orth_proj`` and ``add_SHlight`` are checked against the reference's own functions in
oracle/make_raster_golden.py (goldens in tests/golden/render_pieces.npz).  The rasterisation convention itself is the in-repo
standard_rasterize one (oracle/rasterize_oracle.c), not pytorch3d's -- that part of the reference's render path is
third-party and PARITY UNPINNED."""
import math

import torch
import torch.nn.functional as F

_PI = math.pi
SH_CONST = torch.tensor([1 / math.sqrt(4 * _PI)] + [((2 * _PI) / 3) * math.sqrt(3 / (4 * _PI))] * 3 +
                        [(_PI / 4) * 3 * math.sqrt(5 / (12 * _PI))] * 3 +
                        [(_PI / 4) * (3 / 2) * math.sqrt(5 / (12 * _PI)), (_PI / 4) * (1 / 2) * math.sqrt(5 / (4 * _PI))])


def batch_orth_proj(X, camera):
    camera = camera.reshape(-1, 1, 3)
    return camera[:, :, 0:1] * torch.cat([X[:, :, :2] + camera[:, :, 1:], X[:, :, 2:]], 2)


def vertex_normals(vertices, faces):
    B, V = vertices.shape[:2]
    n = torch.zeros(B, V, 3, dtype=vertices.dtype)
    for b in range(B):
        v = vertices[b][faces]                           # (F,3,3)
        n[b].index_add_(0, faces[:, 1], torch.cross(v[:, 2] - v[:, 1], v[:, 0] - v[:, 1], dim=1))
        n[b].index_add_(0, faces[:, 2], torch.cross(v[:, 0] - v[:, 2], v[:, 1] - v[:, 2], dim=1))
        n[b].index_add_(0, faces[:, 0], torch.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0], dim=1))
    return F.normalize(n, eps=1e-6, dim=2)


def add_sh_light(normal_images, sh_coeff):
    """renderer.py:207-221. normal_images (B,3,H,W), sh_coeff (B,9,3) -> (B,3,H,W)."""
    N = normal_images
    sh = torch.stack([torch.ones_like(N[:, 0]), N[:, 0], N[:, 1], N[:, 2], N[:, 0] * N[:, 1], N[:, 0] * N[:, 2],
                      N[:, 1] * N[:, 2], N[:, 0] ** 2 - N[:, 1] ** 2, 3 * (N[:, 2] ** 2) - 1], 1)
    sh = sh * SH_CONST.to(N.dtype)[None, :, None, None]
    return torch.sum(sh_coeff[:, :, :, None, None] * sh[:, :, None, :, :], 1)


def shade(tri, bary, face_uv, face_normals, albedo, sh):
    """tri (B,H,W) int, bary (B,H,W,3), face_uv (F,3,2), face_normals (B,F,3,3), albedo (B,3,T,T), sh (B,9,3)
    -> images (B,3,H,W), normal_images (B,3,H,W), cond (B,6,H,W)."""
    B, H, W = tri.shape
    mask = tri >= 0
    idx = tri.clamp(min=0).long()
    uv = (bary[..., None] * face_uv[idx]).sum(-2) * mask[..., None]                    # (B,H,W,2); 0 where uncovered
    fn = torch.stack([face_normals[b][idx[b]] for b in range(B)])                       # (B,H,W,3,3)
    nrm = (bary[..., None] * fn).sum(-2) * mask[..., None]
    alb = F.grid_sample(albedo, uv, mode="bilinear", padding_mode="zeros", align_corners=False)
    normal_images = nrm.permute(0, 3, 1, 2)
    images = alb * add_sh_light(normal_images, sh) * mask[:, None].to(alb.dtype)
    tq = torch.floor(images.clamp(0, 255)) / 255.0
    nq = torch.floor(normal_images.clamp(0, 1) * 255) / 255.0
    cond = torch.cat([tq.clamp(0, 1) * 2 - 1, nq.clamp(0, 1) * 2 - 1], 1)
    return images, normal_images, cond
