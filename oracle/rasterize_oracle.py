"""TEST INFRASTRUCTURE ONLY -- Python face of the rasteriser oracles.

* ``oracle_rasterize`` / ``oracle_rasterize_colors``: the C restatement (oracle/rasterize_oracle.c) via ctypes.
* ``ref_rasterize`` / ``ref_rasterize_colors``: the UNMODIFIED reference kernels compiled for the host
  (oracle/_ref/libstandard_rasterize_ref.so, built by oracle/Makefile in the build container).
* ``visibility_pixels`` / ``get_visibility``: restatement of visibility.py:29-60 (NDC -> pixel mapping, buffer
  initialisation, visible-vertex set).
* ``bary_backward_oracle``: the backward oracle (absent in the reference): torch autograd over the
  barycentric / depth / colour formulas (kernel.cu:79-109,:148,:225) on the (pixel, face) pairs the forward chose.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "librasterize_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libstandard_rasterize_ref.so")
_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int32)


def _lib(path, build_target):
    src = os.path.join(_HERE, "rasterize_oracle.c")
    stale = build_target == "oracle" and os.path.isfile(path) and os.path.getmtime(path) < os.path.getmtime(src)
    if not os.path.isfile(path) or stale:
        import subprocess
        subprocess.run(["make", "-C", _HERE, build_target], check=True, capture_output=True)
    return ctypes.CDLL(path)


def have_ref():
    return os.path.isfile(_REF_SO)


def _buffers(batch, h, w, depth, tri, out3):
    depth = np.full((batch, h, w), 1e6, np.float32) if depth is None else np.ascontiguousarray(depth, np.float32)
    tri = np.full((batch, h, w), -1, np.int32) if tri is None else np.ascontiguousarray(tri, np.int32)
    out3 = np.zeros((batch, h, w, 3), np.float32) if out3 is None else np.ascontiguousarray(out3, np.float32)
    return depth, tri, out3


def _run(lib, prefix, fv, colors, h, w, depth, tri, out3):
    fv = np.ascontiguousarray(fv, np.float32)
    batch, ntri = fv.shape[:2]
    depth, tri, out3 = _buffers(batch, h, w, depth, tri, out3)
    args = [fv.ctypes.data_as(_fp)]
    if colors is not None:
        colors = np.ascontiguousarray(colors, np.float32)
        args.append(colors.ctypes.data_as(_fp))
        fn = getattr(lib, prefix + "_standard_rasterize_colors")
    else:
        fn = getattr(lib, prefix + "_standard_rasterize")
    fn.restype = None
    fn(*args, depth.ctypes.data_as(_fp), tri.ctypes.data_as(_ip), out3.ctypes.data_as(_fp),
       ctypes.c_int(batch), ctypes.c_int(ntri), ctypes.c_int(h), ctypes.c_int(w))
    return depth, tri, out3


def oracle_rasterize(fv, h, w, depth=None, tri=None, bary=None):
    return _run(_lib(_ORACLE_SO, "oracle"), "oracle", fv, None, h, w, depth, tri, bary)


def oracle_rasterize_colors(fv, colors, h, w, depth=None, tri=None, images=None):
    return _run(_lib(_ORACLE_SO, "oracle"), "oracle", fv, colors, h, w, depth, tri, images)


def oracle_rasterize_pytorch3d(fv_ndc, h, w):
    """pytorch3d ``rasterize_meshes`` conventions (oracle/rasterize_oracle.c, PARITY UNPINNED): face vertices in NDC (x, y
    already negated like renderer.py:55) -> (zbuf (B,h,w) with -1 where empty, pix_to_face (B,h,w) index within the mesh
    or -1, bary (B,h,w,3) with -1 where empty) -- the values pytorch3d returns for faces_per_pixel = 1."""
    lib = _lib(_ORACLE_SO, "oracle")
    fv = np.ascontiguousarray(fv_ndc, np.float32)
    batch, ntri = fv.shape[:2]
    depth = np.full((batch, h, w), np.inf, np.float32)
    tri = np.full((batch, h, w), -1, np.int32)
    bary = np.full((batch, h, w, 3), -1.0, np.float32)
    lib.oracle_rasterize_pytorch3d.restype = None
    lib.oracle_rasterize_pytorch3d(fv.ctypes.data_as(_fp), depth.ctypes.data_as(_fp), tri.ctypes.data_as(_ip),
                                   bary.ctypes.data_as(_fp), ctypes.c_int(batch), ctypes.c_int(ntri), ctypes.c_int(h),
                                   ctypes.c_int(w))
    depth[tri < 0] = -1.0
    return depth, tri, bary


def ref_rasterize(fv, h, w, depth=None, tri=None, bary=None):
    return _run(ctypes.CDLL(_REF_SO), "ref", fv, None, h, w, depth, tri, bary)


def ref_rasterize_colors(fv, colors, h, w, depth=None, tri=None, images=None):
    return _run(ctypes.CDLL(_REF_SO), "ref", fv, colors, h, w, depth, tri, images)


def visibility_pixels(vertices, h, w):
    """visibility.py:36-40: x*w/2+w/2, y*h/2+h/2, z - min(z over the whole batch) + 1 (fp32)."""
    v = np.array(vertices, np.float32, copy=True)
    v[..., 0] = v[..., 0] * np.float32(w) / np.float32(2) + np.float32(w / 2)
    v[..., 1] = v[..., 1] * np.float32(h) / np.float32(2) + np.float32(h / 2)
    v[..., 2] = v[..., 2] - v[..., 2].min() + np.float32(1)
    return v


def face_vertices(vertices, faces):
    """visibility.py:9-27: gather (B,V,3) by (B,F,3) -> (B,F,3,3)."""
    return np.stack([vertices[b][faces[b]] for b in range(vertices.shape[0])]).astype(np.float32)


def get_visibility(vertices, faces, h, w, rasterize=oracle_rasterize):
    """visibility.py:29-60 -> (B,V) 0/1 per-vertex visibility (vertices of every triangle that owns a pixel)."""
    pix = visibility_pixels(vertices, h, w)
    _, tri, _ = rasterize(face_vertices(pix, faces), h, w)
    vis = np.zeros(vertices.shape[:2], np.float32)
    for b in range(vertices.shape[0]):
        t = np.unique(tri[b])
        t = t[t >= 0]
        vis[b, np.unique(faces[b][t])] = 1.0
    return vis


def load_obj_vf(path, with_uv=False):
    vs, fs, vts, fts = [], [], [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                vs.append([float(x) for x in p[1:]])
            elif p[0] == "vt":
                vts.append([float(x) for x in p[1:3]])
            elif p[0] == "f":
                fs.append([int(x.split("/")[0]) - 1 for x in p[1:4]])
                if with_uv:
                    fts.append([int(x.split("/")[1]) - 1 for x in p[1:4]])
    if with_uv:
        return (np.asarray(vs, np.float64), np.asarray(fs, np.int64), np.asarray(vts, np.float64),
                np.asarray(fts, np.int64))
    return np.asarray(vs, np.float64), np.asarray(fs, np.int64)


# ----------------------------------------------------------------------------- backward oracle
def interp_torch(fv, tri, colors=None):
    """Differentiable re-evaluation at the winners: returns (bary (B,h,w,3), depth (B,h,w), images or None).
    fv (B,F,3,3) torch tensor (requires_grad ok), tri (B,h,w) int tensor from the forward."""
    b, h, w = tri.shape
    mask = tri >= 0
    idx = tri.clamp(min=0).long()
    f = torch.gather(fv.reshape(b, -1, 9), 1, idx.reshape(b, -1, 1).expand(-1, -1, 9)).reshape(b, h, w, 3, 3)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=fv.dtype), torch.arange(w, dtype=fv.dtype), indexing="ij")
    p0, p1, p2 = f[..., 0, :2], f[..., 1, :2], f[..., 2, :2]
    p = torch.stack([xs, ys], -1).expand(b, h, w, 2)
    v0, v1, v2 = p2 - p0, p1 - p0, p - p0
    d00, d01, d02 = (v0 * v0).sum(-1), (v0 * v1).sum(-1), (v0 * v2).sum(-1)
    d11, d12 = (v1 * v1).sum(-1), (v1 * v2).sum(-1)
    den = d00 * d11 - d01 * d01
    inv = torch.where(den == 0, torch.zeros_like(den), 1.0 / torch.where(den == 0, torch.ones_like(den), den))
    u = (d11 * d02 - d01 * d12) * inv
    v = (d00 * d12 - d01 * d02) * inv
    bw = torch.stack([1 - u - v, v, u], -1)
    z = f[..., :, 2]
    zs = torch.where(mask[..., None], z, torch.ones_like(z))
    depth = 1.0 / (bw / zs).sum(-1)
    m = mask.to(fv.dtype)
    out_img = None
    if colors is not None:
        c = torch.gather(colors.reshape(b, -1, 9), 1, idx.reshape(b, -1, 1).expand(-1, -1, 9)).reshape(b, h, w, 3, 3)
        out_img = (bw[..., :, None] * c).sum(-2) * m[..., None]
    return bw * m[..., None], depth * m, out_img
