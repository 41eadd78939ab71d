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


def _forward(face_vertices, face_colors, depth_buffer, triangle_buffer, out3, h, w):
    _check_input("face_vertices", face_vertices, torch.float32)
    _check_input("depth_buffer", depth_buffer, torch.float32)
    _check_input("triangle_buffer", triangle_buffer, torch.int32)
    _check_input("baryw_buffer/images", out3, torch.float32)
    if face_colors is not None:
        _check_input("face_colors", face_colors, torch.float32)
    B, F = face_vertices.shape[:2]
    assert tuple(face_vertices.shape[2:]) == (3, 3)
    assert tuple(depth_buffer.shape) == (B, h, w) and tuple(triangle_buffer.shape) == (B, h, w)
    assert tuple(out3.shape) == (B, h, w, 3)
    nws = lib.gifb200_rasterize_workspace_bytes(B, F, h, w)
    ws = ops._workspace(nws, face_vertices.device)
    check(lib.gifb200_rasterize_fwd(ptr(face_vertices), ptr(face_colors), ptr(depth_buffer), ptr(triangle_buffer),
                                    ptr(out3), B, F, h, w, ptr(ws), nws, stream()), "gifb200_rasterize_fwd")


def standard_rasterize(face_vertices, depth_buffer, triangle_buffer, baryw_buffer, h, w):
    _forward(face_vertices, None, depth_buffer, triangle_buffer, baryw_buffer, h, w)
    return [depth_buffer, triangle_buffer, baryw_buffer]


def standard_rasterize_colors(face_vertices, face_colors, depth_buffer, triangle_buffer, images, h, w):
    _forward(face_vertices, face_colors, depth_buffer, triangle_buffer, images, h, w)
    return [depth_buffer, triangle_buffer, images]


# ---------------------------------------------------------------------------------------------- differentiable form
class _Rasterize(torch.autograd.Function):
    """(face_vertices[, face_colors]) -> (depth, triangle, bary-or-image) on fresh buffers initialised like
    visibility.py:42-44 (depth 1e6, triangle -1, payload 0); backward = gifb200_rasterize_bwd."""

    @staticmethod
    def forward(ctx, face_vertices, face_colors, h, w):
        fv = face_vertices.contiguous()
        fc = None if face_colors is None else face_colors.contiguous()
        B = fv.shape[0]
        depth = torch.full((B, h, w), 1e6, dtype=torch.float32, device=fv.device)
        tri = torch.full((B, h, w), -1, dtype=torch.int32, device=fv.device)
        out3 = torch.zeros((B, h, w, 3), dtype=torch.float32, device=fv.device)
        _forward(fv, fc, depth, tri, out3, h, w)
        ctx.save_for_backward(fv, fc, tri)
        ctx.mark_non_differentiable(tri)
        ctx.hw = (h, w)
        return depth, tri, out3

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_depth, _g_tri, g_out3):
        fv, fc, tri = ctx.saved_tensors
        h, w = ctx.hw
        B, F = fv.shape[:2]
        g_fv = torch.zeros_like(fv)
        g_fc = torch.zeros_like(fc) if fc is not None else None
        g_depth = None if g_depth is None else g_depth.contiguous()
        g_out3 = None if g_out3 is None else g_out3.contiguous()
        check(lib.gifb200_rasterize_bwd(ptr(fv), ptr(fc), ptr(tri), ptr(g_out3) if fc is None else None,
                                        ptr(g_out3) if fc is not None else None, ptr(g_depth), ptr(g_fv), ptr(g_fc),
                                        B, F, h, w, stream()), "gifb200_rasterize_bwd")
        return g_fv, g_fc, None, None


def rasterize(face_vertices, h, w, face_colors=None):
    """Differentiable rasterisation: returns (depth (B,h,w), triangle (B,h,w) int32, bary or colour image (B,h,w,3))."""
    return _Rasterize.apply(face_vertices, face_colors, h, w)


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
