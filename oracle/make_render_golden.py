#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY (build container) -- golden for the COMPOSITE conditioning render.

Runs the reference's UNMODIFIED ``Renderer`` (my_utils/photometric_optimization/renderer.py:87-305: forward, add_SHlight,
render_normal, and Pytorch3dRasterizer.forward's attribute interpolation :69-84) exactly as
``gif_helper.render_utils.render_tex_and_normal`` (gif_helper.py:24-40) and ``OverLayViz.get_rendered_mesh``
(visualize_flame_overlay.py:29-31) drive it, on synthetic FLAME-shaped inputs.  The three names the reference imports from
pytorch3d (absent: an unpinned third-party fork) are supplied at the module boundary:
    rasterize_meshes -> oracle/rasterize_oracle.c's restatement of pytorch3d's published rules (PARITY UNPINNED for that step),
    Meshes           -> a two-field holder (verts, faces),
    load_obj         -> a parser of the reference's own head_template_mesh.obj returning pytorch3d's (verts, faces, aux) triple.
Everything downstream of (pix_to_face, bary) is therefore the reference's own code.  Output: tests/golden/render_composite.npz
(inputs + textured image, normal image, alpha, quantised maps), consumed by tests/test_render_oracle.py (pins
oracle/render_oracle.shade to it) and tests/test_render_gpu.py (the CUDA path)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import golden_util as gu  # noqa: E402
from oracle import rasterize_oracle as RO  # noqa: E402
from oracle import ref_import  # noqa: E402


class Meshes:
    def __init__(self, verts, faces):
        self.verts, self.faces = verts, faces


def rasterize_meshes(meshes, image_size, blur_radius, faces_per_pixel, bin_size, max_faces_per_bin, perspective_correct):
    assert blur_radius == 0.0 and faces_per_pixel == 1 and not perspective_correct
    B, F = meshes.faces.shape[:2]
    fv = torch.stack([meshes.verts[b][meshes.faces[b]] for b in range(B)]).numpy()
    zbuf, tri, bary = RO.oracle_rasterize_pytorch3d(fv, image_size, image_size)
    p2f = np.where(tri >= 0, tri.astype(np.int64) + np.arange(B)[:, None, None] * F, -1)        # packed face index
    return (torch.from_numpy(p2f)[..., None], torch.from_numpy(zbuf)[..., None], torch.from_numpy(bary)[:, :, :, None, :],
            torch.full((B, image_size, image_size, 1), -1.0))


def load_obj(filename):
    v, f, vt, ft = RO.load_obj_vf(filename, with_uv=True)
    return (torch.from_numpy(v).float(), types.SimpleNamespace(verts_idx=torch.from_numpy(f), textures_idx=torch.from_numpy(ft)),
            types.SimpleNamespace(verts_uvs=torch.from_numpy(vt).float()))


def main():
    from gif_b200.flame_synth import synthetic_flame_params        # host-side synthetic inputs (no kernels involved)
    ref_import.load()
    from my_utils.photometric_optimization import renderer as ren, util as ref_util
    ren.Meshes, ren.rasterize_meshes, ren.load_obj = Meshes, rasterize_meshes, load_obj
    S, B = 128, 2
    obj = os.path.join(ref_import.REF_ROOT, "my_utils", "photometric_optimization", "data", "head_template_mesh.obj")
    R = ren.Renderer(S, obj_filename=obj)
    verts, cam, alb, lights = synthetic_flame_params(B, seed=7)
    # gif_helper.py:24-40
    trans = ref_util.batch_orth_proj(verts, cam)
    trans[:, :, 1:] = -trans[:, :, 1:]
    with torch.no_grad():
        res = R(verts, trans, alb, lights=lights)                   # mutates trans (z + 10), as in the reference
        textured, normals = res["images"], res["normals"]
        normal_images = R.render_normal(trans, normals)
    # visualize_flame_overlay.py:29-31
    tq = torch.floor(textured.clamp(0, 255)) / 255.0
    nq = torch.floor(normal_images.clamp(0, 1) * 255) / 255.0
    out = dict(verts=verts.numpy(), cam=cam.numpy(), albedo=alb.numpy(), lights=lights.numpy(), images=textured.numpy(),
               normal_images=normal_images.numpy(), alpha=res["alpha_images"].numpy(), tex_quantised=tq.numpy(), normal_quantised=nq.numpy())
    np.savez_compressed(os.path.join(gu.GOLDEN_DIR, "render_composite.npz"), **out)
    print("render_composite.npz: coverage", float(res["alpha_images"].mean()), "images max", float(textured.max()))


if __name__ == "__main__":
    main()
