"""Rasteriser with the interface of the reference's pybind module
(my_utils/standard_rasterize_cuda/standard_rasterize_cuda.cpp:26-40,59-75,79-82) and of visibility.py, plus the
differentiable wrapper the reference lacks.

``standard_rasterize(face_vertices, depth_buffer, triangle_buffer, baryw_buffer, h, w)`` and
``standard_rasterize_colors(face_vertices, face_colors, depth_buffer, triangle_buffer, images, h, w)`` mutate the
caller's buffers in place and return them, exactly as the reference does; all tensors must be CUDA + contiguous
(the reference's CHECK_INPUT, .cpp:21-23) or a RuntimeError is raised.
"""
import torch

from . import ops
from ._lib import GifB200Error, check, lib, ptr, stream


def _check_input(name, t, dtype):
    if not t.is_cuda:
        raise GifB200Error(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise GifB200Error(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise GifB200Error(f"{name} must be {dtype}")


CONVENTIONS = {"standard": 0, "pytorch3d": 1}


def _forward(face_vertices, face_colors, depth_buffer, triangle_buffer, out3, h, w, face_colors2=None, out3b=None,
             convention="standard"):
    _check_input("face_vertices", face_vertices, torch.float32)
    _check_input("depth_buffer", depth_buffer, torch.float32)
    _check_input("triangle_buffer", triangle_buffer, torch.int32)
    _check_input("baryw_buffer/images", out3, torch.float32)
    if face_colors is not None:
        _check_input("face_colors", face_colors, torch.float32)
    if face_colors2 is not None:
        _check_input("face_colors2", face_colors2, torch.float32)
        _check_input("images2", out3b, torch.float32)
    B, F = face_vertices.shape[:2]
    assert tuple(face_vertices.shape[2:]) == (3, 3)
    assert tuple(depth_buffer.shape) == (B, h, w) and tuple(triangle_buffer.shape) == (B, h, w)
    assert tuple(out3.shape) == (B, h, w, 3)
    nws = lib.gifb200_rasterize_workspace_bytes(B, F, h, w)
    ws = ops._workspace(nws, face_vertices.device)
    check(lib.gifb200_rasterize_fwd_ex(ptr(face_vertices), ptr(face_colors), ptr(face_colors2), ptr(depth_buffer),
                                       ptr(triangle_buffer), ptr(out3), ptr(out3b), B, F, h, w, CONVENTIONS[convention],
                                       ptr(ws), nws, stream()), "gifb200_rasterize_fwd_ex")


def standard_rasterize(face_vertices, depth_buffer, triangle_buffer, baryw_buffer, h, w):
    _forward(face_vertices, None, depth_buffer, triangle_buffer, baryw_buffer, h, w)
    return [depth_buffer, triangle_buffer, baryw_buffer]


def standard_rasterize_colors(face_vertices, face_colors, depth_buffer, triangle_buffer, images, h, w):
    _forward(face_vertices, face_colors, depth_buffer, triangle_buffer, images, h, w)
    return [depth_buffer, triangle_buffer, images]


# ---------------------------------------------------------------------------------------------- differentiable form
class _Rasterize(torch.autograd.Function):
    """(face_vertices[, face_colors[, face_colors2]]) -> (depth, triangle, bary-or-image[, image2]) on fresh buffers
    initialised like visibility.py:42-44 (depth 1e6, triangle -1, payload 0) for the in-repo convention, or like pytorch3d
    (zbuf -1, pix_to_face -1, bary -1 where empty) for convention "pytorch3d"; backward = gifb200_rasterize_bwd_ex (a
    per-face gather: outputs are written, not accumulated)."""

    @staticmethod
    def forward(ctx, face_vertices, face_colors, face_colors2, h, w, convention):
        fv = face_vertices.contiguous()
        fc = None if face_colors is None else face_colors.contiguous()
        fc2 = None if face_colors2 is None else face_colors2.contiguous()
        B = fv.shape[0]
        p3d = convention == "pytorch3d"
        depth = torch.full((B, h, w), float("inf") if p3d else 1e6, dtype=torch.float32, device=fv.device)
        tri = torch.full((B, h, w), -1, dtype=torch.int32, device=fv.device)
        out3 = torch.full((B, h, w, 3), -1.0 if (p3d and fc is None) else 0.0, dtype=torch.float32, device=fv.device)
        out3b = torch.zeros((B, h, w, 3), dtype=torch.float32, device=fv.device) if fc2 is not None else None
        _forward(fv, fc, depth, tri, out3, h, w, fc2, out3b, convention)
        if p3d:
            depth.masked_fill_(tri < 0, -1.0)
        ctx.save_for_backward(fv, fc, fc2, tri)
        ctx.mark_non_differentiable(tri)
        ctx.set_materialize_grads(False)          # an unused output (e.g. depth) arrives as None, not as a tensor of zeros
        ctx.cfg = (h, w, convention)
        return (depth, tri, out3) if fc2 is None else (depth, tri, out3, out3b)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_depth, _g_tri, g_out3, g_out3b=None):
        fv, fc, fc2, tri = ctx.saved_tensors
        h, w, convention = ctx.cfg
        B, F = fv.shape[:2]
        g_fv = torch.empty_like(fv)
        g_fc = torch.empty_like(fc) if fc is not None else None
        g_fc2 = torch.empty_like(fc2) if fc2 is not None else None
        g_depth = None if g_depth is None else g_depth.contiguous()
        g_out3 = None if g_out3 is None else g_out3.contiguous()
        g_out3b = None if g_out3b is None else g_out3b.contiguous()
        check(lib.gifb200_rasterize_bwd_ex(ptr(fv), ptr(fc), ptr(fc2), ptr(tri), ptr(g_out3) if fc is None else None,
                                           ptr(g_out3) if fc is not None else None, ptr(g_out3b) if fc2 is not None else None,
                                           ptr(g_depth), ptr(g_fv), ptr(g_fc), ptr(g_fc2), B, F, h, w, CONVENTIONS[convention],
                                           stream()), "gifb200_rasterize_bwd_ex")
        return g_fv, g_fc, g_fc2, None, None, None


def rasterize(face_vertices, h, w, face_colors=None, face_colors2=None, convention="standard"):
    """Differentiable rasterisation: returns (depth (B,h,w), triangle (B,h,w) int32, bary or colour image (B,h,w,3)[, second
    attribute image (B,h,w,3)]).  ``face_colors2``: a second per-corner attribute set interpolated from the SAME
    rasterisation (texture + normal render = one pass).  ``convention``: "standard" (the in-repo standard_rasterize, pixel
    space) or "pytorch3d" (NDC input, x/y already negated, renderer.py:46-67; see include/gifb200.h)."""
    return _Rasterize.apply(face_vertices, face_colors, face_colors2, h, w, convention)


def rasterize_meshes(face_vertices_ndc, image_size):
    """What the reference gets back from pytorch3d.renderer.mesh.rasterize_meshes(meshes, image_size, blur_radius=0,
    faces_per_pixel=1, perspective_correct=False) (renderer.py:59-67), for a batch of same-topology meshes given as
    face vertices (B,F,3,3) in NDC: (pix_to_face (B,S,S,1) int64 with the packed offset b*F, -1 where empty;
    zbuf (B,S,S,1); bary_coords (B,S,S,1,3); dists None -- the reference never reads it)."""
    zbuf, tri, bary = rasterize(face_vertices_ndc, image_size, image_size, convention="pytorch3d")
    F = face_vertices_ndc.shape[1]
    off = torch.arange(face_vertices_ndc.shape[0], device=tri.device, dtype=torch.int64)[:, None, None] * F
    pix_to_face = torch.where(tri >= 0, tri.long() + off, torch.full_like(tri, -1, dtype=torch.int64))
    return pix_to_face[..., None], zbuf[..., None], bary[:, :, :, None, :], None


class Pytorch3dRasterizer(torch.nn.Module):
    """photometric_optimization/renderer.py:19-84 on the gif_b200 rasteriser: same constructor, same forward
    (vertices (B,V,3) projected to NDC, faces (B,F,3), attributes (B,F,3,D)) -> (B, D+1, S, S) interpolated attributes +
    visibility mask."""

    def __init__(self, image_size=224):
        super().__init__()
        self.image_size = image_size

    def forward(self, vertices, faces, attributes=None):
        fixed = vertices.clone().float()
        fixed[..., :2] = -fixed[..., :2]                                          # renderer.py:54-55
        fv = face_vertices(fixed, faces.int()).contiguous()
        pix_to_face, _zbuf, bary, _ = rasterize_meshes(fv, self.image_size)
        vismask = (pix_to_face > -1).float()                                      # renderer.py:69-84 from here on
        D = attributes.shape[-1]
        attributes = attributes.reshape(attributes.shape[0] * attributes.shape[1], 3, D)
        N, H, W, K, _ = bary.shape
        mask = pix_to_face == -1
        idx = pix_to_face.clamp(min=0).view(N * H * W * K, 1, 1).expand(N * H * W * K, 3, D)
        vals = attributes.gather(0, idx).view(N, H, W, K, 3, D)
        pixel_vals = (bary[..., None] * vals).sum(dim=-2)
        pixel_vals[mask] = 0
        pixel_vals = pixel_vals[:, :, :, 0].permute(0, 3, 1, 2)
        return torch.cat([pixel_vals, vismask[:, :, :, 0][:, None, :, :]], dim=1)


def face_vertices(vertices, faces):
    """visibility.py:9-27."""
    bs, nv = vertices.shape[:2]
    faces = faces + (torch.arange(bs, dtype=torch.int32, device=vertices.device) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[faces.long()]


def get_visibility(vertices, triangles, h, w):
    """visibility.py:29-60 -> (B,V) per-vertex visibility."""
    bz = vertices.shape[0]
    device = vertices.device
    vertices = vertices.clone()
    vertices[..., 0] = vertices[..., 0] * w / 2 + w / 2
    vertices[..., 1] = vertices[..., 1] * h / 2 + h / 2
    vertices[..., 2] = vertices[..., 2] - vertices[..., 2].min() + 1
    depth_buffer = torch.zeros([bz, h, w], device=device).float() + 1e6
    triangle_buffer = torch.zeros([bz, h, w], device=device).int() - 1
    baryw_buffer = torch.zeros([bz, h, w, 3], device=device).float()
    vert_vis = torch.zeros([bz, vertices.shape[1]], device=device).float()
    f_vs = face_vertices(vertices, triangles).contiguous()
    standard_rasterize(f_vs, depth_buffer, triangle_buffer, baryw_buffer, h, w)
    triangle_buffer = triangle_buffer.reshape(bz, -1)
    for i in range(bz):
        tri_visind = torch.unique(triangle_buffer[i])
        tri_visind = tri_visind[tri_visind >= 0].long()
        vert_visind = triangles[i, tri_visind, :].flatten()
        vert_vis[i, torch.unique(vert_visind.long())] = 1.0
    return vert_vis
