"""FLAME conditioning render on the GPU: (vertices, camera-projected vertices, albedo, SH lights) -> textured image,
normal image and the 6-channel condition map the generator consumes.

Follows the arithmetic of the reference's render path
    gif_helper.render_utils.render_tex_and_normal   (my_utils/photometric_optimization/gif_helper.py:24-40)
    Renderer.forward / add_SHlight / render_normal  (my_utils/photometric_optimization/renderer.py:130-221,291-305)
    util.vertex_normals / face_vertices / batch_orth_proj (my_utils/photometric_optimization/util.py:73-83,135-189)
    OverLayViz.get_rendered_mesh quantisation       (my_utils/visualize_flame_overlay.py:29-31)
Two rasterisation conventions (``FlameRenderer(..., convention=...)``), the choice stated, not hidden:
  "pytorch3d" (default) -- what the reference's conditioning maps are actually made with: pytorch3d ``rasterize_meshes`` as
               ``Pytorch3dRasterizer.forward`` calls it (renderer.py:46-67: NDC with x, y negated, pixel centres at +0.5,
               no back-face culling, linear depth, strictly-inside test), gifb200_rasterize_fwd_ex convention 1.  The fork the
               reference pins is absent and unversioned: PARITY UNPINNED, checked against the restatement of
               the published rules in oracle/rasterize_oracle.c;
  "standard"  -- the in-repo ``standard_rasterize`` semantics (gif_b200.rasterize: pixel centres at integer coordinates
               after the visibility.py:38-40 mapping, front faces only, perspective-interpolated depth), bit-exact against
               the reference's own kernels.
Everything downstream of the (triangle, bary) buffers is the reference's formulae, fused into one kernel
(gifb200_render_shade).

Conditioning maps are *data* for the GAN (the reference detaches every attribute, renderer.py:149-150), so this path is
forward-only; the differentiable rasteriser itself is gif_b200.rasterize.rasterize.
"""
import torch
import torch.nn.functional as F

from . import rasterize
from ._lib import check, lib, ptr, stream


def batch_orth_proj(X, camera):
    """util.py:73-83: (s, tx, ty) weak-perspective camera."""
    camera = camera.reshape(-1, 1, 3)
    X_trans = torch.cat([X[:, :, :2] + camera[:, :, 1:], X[:, :, 2:]], 2)
    return camera[:, :, 0:1] * X_trans


def face_vertices(vertices, faces):
    """util.py:135-153: (B,V,C) gathered by (F,3) -> (B,F,3,C)."""
    return vertices[:, faces]


def vertex_normals(vertices, faces):
    """util.py:156-189: area-weighted vertex normals (three index_add of face cross products), normalised (eps 1e-6)."""
    B, V = vertices.shape[:2]
    vf = vertices[:, faces]                                                    # (B,F,3,3)
    n = torch.zeros(B, V, 3, device=vertices.device, dtype=vertices.dtype)
    idx = faces.to(vertices.device)
    n.index_add_(1, idx[:, 1], torch.cross(vf[:, :, 2] - vf[:, :, 1], vf[:, :, 0] - vf[:, :, 1], dim=-1))
    n.index_add_(1, idx[:, 2], torch.cross(vf[:, :, 0] - vf[:, :, 2], vf[:, :, 1] - vf[:, :, 2], dim=-1))
    n.index_add_(1, idx[:, 0], torch.cross(vf[:, :, 1] - vf[:, :, 0], vf[:, :, 2] - vf[:, :, 0], dim=-1))
    return F.normalize(n, eps=1e-6, dim=2)


class FlameRenderer(torch.nn.Module):
    """Renderer (renderer.py:87-127) for a fixed topology: faces (F,3), per-corner UVs from (uvcoords (Vt,2), uvfaces (F,3))."""

    def __init__(self, faces, uvcoords, uvfaces, image_size=256, convention="pytorch3d"):
        super().__init__()
        if convention not in rasterize.CONVENTIONS:
            raise ValueError(f"convention must be one of {sorted(rasterize.CONVENTIONS)}")
        self.convention = convention
        self.image_size = image_size
        self.register_buffer("faces", faces.long())
        uv = torch.cat([uvcoords, torch.ones_like(uvcoords[:, :1])], -1) * 2 - 1       # renderer.py:107-109
        uv[:, 1] = -uv[:, 1]
        self.register_buffer("face_uv", uv[uvfaces.long()][:, :, :2].contiguous().float())   # (F,3,2) grid coordinates

    @torch.no_grad()
    def forward(self, vertices, transformed_vertices, albedos, lights, want_cond=True):
        """vertices (B,V,3) world space; transformed_vertices (B,V,3) projected to [-1,1] (x right, y down after the flip
        of gif_helper.py:27); albedos (B,3,T,T); lights (B,9,3).  Returns dict(images (B,3,H,W), normal_images (B,3,H,W),
        alpha (B,1,H,W), cond (B,6,H,W) in [-1,1], triangle (B,H,W))."""
        B = vertices.shape[0]
        H = W = self.image_size
        tv = transformed_vertices.clone().float()
        tv[:, :, 2] = tv[:, :, 2] + 10                                             # renderer.py:139
        pix = tv.clone()
        if self.convention == "pytorch3d":
            pix[..., :2] = -pix[..., :2]                                           # renderer.py:54-55, NDC in
            depth0 = float("inf")
        else:                                                                      # visibility.py:38-40 pixel mapping
            pix[..., 0] = tv[..., 0] * W / 2 + W / 2
            pix[..., 1] = tv[..., 1] * H / 2 + H / 2
            pix[..., 2] = tv[..., 2] - tv[..., 2].min() + 1
            depth0 = 1e6
        fv = face_vertices(pix, self.faces).contiguous()
        normals = vertex_normals(vertices.float(), self.faces)                     # renderer.py:143
        fn = face_vertices(normals, self.faces).contiguous()
        depth = torch.full((B, H, W), depth0, device=fv.device)
        tri = torch.full((B, H, W), -1, dtype=torch.int32, device=fv.device)
        bary = torch.zeros((B, H, W, 3), device=fv.device)
        rasterize._forward(fv, None, depth, tri, bary, H, W, convention=self.convention)
        tex = torch.empty((B, H, W, 3), device=fv.device)
        nrm = torch.empty((B, H, W, 3), device=fv.device)
        cond = torch.empty((B, H, W, 6), device=fv.device) if want_cond else None
        alb = albedos.contiguous().float()
        sh = lights.contiguous().float()
        check(lib.gifb200_render_shade(ptr(tri), ptr(bary), ptr(self.face_uv), ptr(fn), ptr(alb), ptr(sh), ptr(tex), ptr(nrm),
                                       ptr(cond), B, self.faces.shape[0], H, W, alb.shape[-1], stream()),
              "gifb200_render_shade")
        return {"images": tex.permute(0, 3, 1, 2), "normal_images": nrm.permute(0, 3, 1, 2),
                "alpha": (tri >= 0).float()[:, None], "cond": None if cond is None else cond.permute(0, 3, 1, 2),
                "triangle": tri, "normals": normals}

    def render_tex_and_normal(self, verts, cam, albedos, lights):
        """gif_helper.py:24-40 for given world vertices: orthographic projection, y/z flip, render."""
        trans = batch_orth_proj(verts, cam)
        trans[:, :, 1:] = -trans[:, :, 1:]
        out = self.forward(verts, trans, albedos, lights)
        return out["images"], out["normal_images"], out["cond"]
