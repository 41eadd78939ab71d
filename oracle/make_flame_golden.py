#!/usr/bin/env python
"""Container-only: pins oracle/flame_oracle.py to the *unmodified* reference decoder functions
(my_utils/photometric_optimization/models/lbs.py, FLAME.py) on the synthetic FLAME-shaped model and writes
tests/golden/flame_lbs.npz (reference outputs, fp32 run and fp64 run) + appends the agreement to
tests/golden/ORACLE_VS_REFERENCE.txt.   usage: python oracle/make_flame_golden.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flame_oracle as FO  # noqa: E402
from oracle import ref_import  # noqa: E402
from gif_b200.flame_synth import synthetic_flame_model  # noqa: E402


def params(B, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    shape = torch.randn(B, 100, generator=g)
    exp = torch.randn(B, 50, generator=g)
    pose = torch.cat([(torch.rand(B, 3, generator=g) * 2 - 1) * torch.tensor([0.3, 0.9, 0.2]),      # global rotation
                      torch.rand(B, 1, generator=g) * 0.5, (torch.rand(B, 2, generator=g) * 2 - 1) * 0.05], 1)  # jaw
    eye = (torch.rand(B, 6, generator=g) * 2 - 1) * 0.3
    neck = (torch.rand(B, 3, generator=g) * 2 - 1) * 0.3
    pose[0] = 0; eye[0] = 0; neck[0] = 0          # rest pose (exercises the 1e-8 epsilon of batch_rodrigues)
    return [t.to(dtype) for t in (shape, exp, pose, eye, neck)]


def reference_forward(ref_lbs, ref_flame, m, dt, shape, exp, pose, eye, neck):
    """FLAME.forward (FLAME.py:175-216) spelled with the reference's own functions (the class itself needs the pickle)."""
    B = shape.shape[0]
    betas = torch.cat([shape, exp], 1)
    full_pose = torch.cat([pose[:, :3], neck, pose[:, 3:], eye], 1)
    f = lambda k: m[k].to(dt)
    verts, joints = ref_lbs.lbs(betas, full_pose, f("v_template")[None].expand(B, -1, -1), f("shapedirs"), f("posedirs"),
                                f("J_regressor"), m["parents"], f("lbs_weights"), dtype=dt)
    chain = FO.neck_kin_chain(m["parents"])
    d_idx, d_bc = ref_flame.FLAME._find_dynamic_lmk_idx_and_bcoords(None, full_pose, m["dynamic_lmk_faces_idx"],
                                                                    f("dynamic_lmk_bary_coords"), chain, dtype=dt)
    idx = torch.cat([d_idx, m["lmk_faces_idx"][None].expand(B, -1)], 1)
    bc = torch.cat([d_bc, f("lmk_bary_coords")[None].expand(B, -1, -1)], 1)
    l2 = ref_lbs.vertices2landmarks(verts, m["faces"], idx, bc)
    l3 = ref_lbs.vertices2landmarks(verts, m["faces"], m["full_lmk_faces_idx"].repeat(B, 1), f("full_lmk_bary_coords").repeat(B, 1, 1))
    return verts, joints, l2, l3


def main():
    ref_import.load()
    ref_lbs = importlib.import_module("my_utils.photometric_optimization.models.lbs")
    ref_flame = importlib.import_module("my_utils.photometric_optimization.models.FLAME")
    # torch 2.11's einsum hands back a permuted view where the reference (written for torch 1.x) expects a contiguous
    # tensor (lbs.py:351-353 `.view`): same values, made contiguous -- the only accommodation, no arithmetic is touched.
    _v2j = ref_lbs.vertices2joints
    ref_lbs.vertices2joints = lambda jr, v: _v2j(jr, v).contiguous()
    m = synthetic_flame_model()
    out, lines = {}, []
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        p = params(4, 2024, dt)
        rv, rj, r2, r3 = reference_forward(ref_lbs, ref_flame, m, dt, *p)
        ov, o2, o3 = FO.flame_forward(m, p[0], p[1], p[2], p[3], p[4])
        _, oj = FO.lbs(torch.cat([p[0], p[1]], 1), torch.cat([p[2][:, :3], p[4], p[2][:, 3:], p[3]], 1),
                       m["v_template"].to(dt), m["shapedirs"].to(dt), m["posedirs"].to(dt), m["J_regressor"].to(dt),
                       m["parents"], m["lbs_weights"].to(dt))
        for name, a, b in (("vertices", ov, rv), ("joints", oj, rj), ("landmarks2d", o2, r2), ("landmarks3d", o3, r3)):
            err = float((a - b).abs().max())
            tol = 1e-12 if dt == torch.float64 else 2e-6
            assert err < tol, (tag, name, err)
            lines.append(f"flame_{name}[{tag}] max|oracle - reference| = {err:.3e}")
            out[f"{name}_{tag}"] = b.numpy().astype(np.float64 if dt == torch.float64 else np.float32)
        if dt == torch.float32:
            for k, v in zip(("shape", "exp", "pose", "eye", "neck"), p):
                out[k] = v.numpy()
    # keep the fixture small: fp64 vertices only for a strided subset (fp32 vertices are complete)
    out["vertices_f64_stride"] = np.int64(7)
    out["vertices_f64"] = out["vertices_f64"][:, ::7]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "flame_lbs.npz"), **out)
    with open(os.path.join(ROOT, "tests", "golden", "ORACLE_VS_REFERENCE.txt"), "a") as fh:
        fh.write("\n# oracle/flame_oracle.py vs reference lbs.py / FLAME.py (synthetic FLAME-shaped model, make_flame_golden.py)\n")
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    texture_golden(ref_import.load(), m)


def texture_golden(R, m):
    """FlameTextureSpace.compute_texture_map (model/stg2_generator.py:376-421), called UNBOUND on a stand-in object that
    carries only the table attributes the method reads (the constructor needs the licence-gated FLAME pickle)."""
    import types
    from oracle import texture_oracle as TO
    from gif_b200.flame_synth import synthetic_texture_data
    td = synthetic_texture_data()
    lines, out = [], {}
    for dt, tag in ((torch.float32, "f32"),):          # the reference method is fp32-only (texture_grid is created float32, :403)
        p = params(3, 77, dt)
        verts, _, _ = FO.flame_forward(m, p[0], p[1], p[2], p[3], p[4])
        g = torch.Generator().manual_seed(5)
        cam = torch.cat([torch.rand(3, 1, generator=g) * 3 + 6, (torch.rand(3, 2, generator=g) * 2 - 1) * 0.03], 1).to(dt)
        src = (torch.rand(3, 3, 48, 40, generator=g) * 2 - 1).to(dt)
        tv = TO.batch_orth_proj(verts, cam).clone()
        tv[:, :, 1:] = -tv[:, :, 1:]
        vn_ref = R.gen.mesh_and_3d_helpers.vertex_normals(tv.float(), m["faces"][None].expand(3, -1, -1)).to(dt) \
            if dt == torch.float32 else TO.vertex_normals(tv, m["faces"])      # the reference helper is fp32-only (:19)
        stand_in = types.SimpleNamespace(
            x_coords=td["x_coords"], y_coords=td["y_coords"], valid_pixel_ids=td["valid_pixel_ids"],
            valid_pixel_3d_faces=torch.from_numpy(td["valid_pixel_3d_faces"]),
            valid_pixel_b_coords=torch.from_numpy(td["valid_pixel_b_coords"]).to(dt))
        with __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            tex_r, mask_r = R.gen.FlameTextureSpace.compute_texture_map(stand_in, src, verts, vn_ref, camera_params=cam)
        tex_o, mask_o = TO.compute_texture_map(src, verts, TO.vertex_normals(tv, m["faces"]), cam, td)
        err = float((tex_o - tex_r).abs().max())
        mism = int((mask_o != mask_r).sum())
        assert err < (1e-12 if dt == torch.float64 else 2e-6) and mism == 0, (tag, err, mism)
        lines.append(f"texture_steal[{tag}] max|oracle - reference| = {err:.3e}, visibility-mask mismatches = {mism}")
        if dt == torch.float32:
            out.update(shape=p[0].numpy(), exp=p[1].numpy(), pose=p[2].numpy(), eye=p[3].numpy(), neck=p[4].numpy(),
                       cam=cam.numpy(), src=src.numpy(), verts=verts.numpy(),
                       tex_sub=tex_r.numpy()[:, :, ::3, ::3], mask=np.packbits(mask_r.numpy()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "texture_steal.npz"), **out)
    with open(os.path.join(ROOT, "tests", "golden", "ORACLE_VS_REFERENCE.txt"), "a") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
