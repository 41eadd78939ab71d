"""TEST INFRASTRUCTURE ONLY -- rasteriser fixtures (run in the build container: python -m oracle.make_raster_golden).

* tests/golden/body_visibility.npz : the reference's only golden artefact for its native kernel --
  data/obj/body.obj -> data/obj/body_vis.obj (demo_vert_visibility.py:12-22: vertices*0.8, get_visibility 512x512);
  the per-vertex 0/1 colours of body_vis.obj are the expected output.
* tests/golden/flame_template.npz  : topology (V=5023, F=9976) + template vertices of
  my_utils/photometric_optimization/data/head_template_mesh.obj (BASELINE config 4 workload).
* tests/golden/raster_cases.npz    : outputs of the UNMODIFIED reference kernels executed on the CPU (oracle/_ref)
  for seeded random meshes, incl. edge cases; the C restatement is asserted bit-identical (modulo exact-zp ties).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402
from oracle import rasterize_oracle as RO  # noqa: E402
import raster_cases  # noqa: E402  (tests/raster_cases.py: seeded mesh generators shared with the tests)

REF = "/root/reference"


def body():
    base = REF + "/my_utils/standard_rasterize_cuda/data/obj/"
    v, f = RO.load_obj_vf(base + "body.obj")
    vv, _ = RO.load_obj_vf(base + "body_vis.obj")
    verts = (v[None, :, :3] * 0.8).astype(np.float32)
    assert np.abs(verts[0] - vv[:, :3]).max() < 1e-6
    gold = vv[:, 3].astype(np.uint8)
    for name, fn in (("oracle", RO.oracle_rasterize), ("reference-on-cpu", RO.ref_rasterize)):
        vis = RO.get_visibility(verts, f[None], 512, 512, rasterize=fn)
        mism = int((vis[0] != gold).sum())
        print(f"body_vis.obj: {name}: {mism} / {gold.size} vertices differ from the checked-in golden")
        assert mism == 0
    np.savez_compressed(os.path.join(gu.GOLDEN_DIR, "body_visibility.npz"), vertices=verts[0],
                        faces=f.astype(np.int32), visible=gold)


def flame():
    v, f, vt, ft = RO.load_obj_vf(REF + "/my_utils/photometric_optimization/data/head_template_mesh.obj", with_uv=True)
    assert v.shape == (5023, 3) and f.shape == (9976, 3) and vt.shape == (5118, 2) and ft.shape == (9976, 3)
    np.savez_compressed(os.path.join(gu.GOLDEN_DIR, "flame_template.npz"), vertices=v.astype(np.float32),
                        faces=f.astype(np.int32), uvcoords=vt.astype(np.float32), uvfaces=ft.astype(np.int32))


def cases():
    out = {}
    for name, (fv, colors, h, w) in raster_cases.all_cases().items():
        d_r, t_r, o_r = (RO.ref_rasterize_colors(fv, colors, h, w) if colors is not None else RO.ref_rasterize(fv, h, w))
        d_o, t_o, o_o = (RO.oracle_rasterize_colors(fv, colors, h, w) if colors is not None
                         else RO.oracle_rasterize(fv, h, w))
        assert np.array_equal(d_r, d_o), name                      # depth is tie-independent: bit-exact
        diff = t_r != t_o
        # the sequential CPU execution of the reference resolves exact-zp ties to the highest index, the oracle to
        # the lowest: any index difference must be such a tie.
        print(f"{name}: covered {int((t_o >= 0).sum())} px, index differences (exact-zp ties) {int(diff.sum())}")
        assert np.array_equal(o_r[~diff], o_o[~diff]), name
        out[name + "_depth"] = d_r
        out[name + "_tri"] = t_o                                    # lowest-index tie-break (the build's policy)
        out[name + "_tie"] = diff
        out[name + "_out3"] = o_o
    np.savez_compressed(os.path.join(gu.GOLDEN_DIR, "raster_cases.npz"), **out)


def render_pieces():
    """Pins oracle/render_oracle.py's vertex_normals / batch_orth_proj / add_sh_light to the reference's own functions
    (my_utils/photometric_optimization/util.py:73-83,156-189; renderer.py:207-221, called unbound on a stand-in object that
    only carries the constant_factor buffer of renderer.py:119-126)."""
    import types
    import torch
    from oracle import ref_import, render_oracle as RD
    ref_import.load()
    from my_utils.photometric_optimization import renderer as ref_renderer, util as ref_util
    z = np.load(os.path.join(gu.GOLDEN_DIR, "flame_template.npz"))
    faces = torch.from_numpy(z["faces"].astype(np.int64))
    g = torch.Generator().manual_seed(5)
    verts = torch.from_numpy(z["vertices"])[None].repeat(2, 1, 1) + 0.003 * torch.randn(2, 5023, 3, generator=g)
    cam = torch.tensor([[8.0, 0.01, -0.02], [9.5, -0.015, 0.005]])
    n_ref = ref_util.vertex_normals(verts, faces[None].expand(2, -1, -1))
    n_or = RD.vertex_normals(verts, faces)
    assert (n_ref - n_or).abs().max() < 1e-6
    p_ref = ref_util.batch_orth_proj(verts, cam)
    assert (p_ref - RD.batch_orth_proj(verts, cam)).abs().max() < 1e-6
    nimg = torch.randn(2, 3, 8, 8, generator=g)
    shc = torch.randn(2, 9, 3, generator=g)
    fake_self = types.SimpleNamespace(constant_factor=RD.SH_CONST.clone())
    s_ref = ref_renderer.Renderer.add_SHlight(fake_self, nimg, shc)
    assert (s_ref - RD.add_sh_light(nimg, shc)).abs().max() < 1e-5
    np.savez_compressed(os.path.join(gu.GOLDEN_DIR, "render_pieces.npz"), verts=verts.numpy(), cam=cam.numpy(),
                        normals=n_ref.numpy(), proj=p_ref.numpy(), nimg=nimg.numpy(), sh=shc.numpy(), shading=s_ref.numpy())
    print("render pieces: vertex_normals / batch_orth_proj / add_SHlight match the reference")


if __name__ == "__main__":
    body()
    flame()
    cases()
    render_pieces()
