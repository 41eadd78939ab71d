"""CPU: the FLAME decoder oracle (oracle/flame_oracle.py) against the goldens produced by the unmodified reference
functions (oracle/make_flame_golden.py -> tests/golden/flame_lbs.npz), plus properties of linear blend skinning."""
import numpy as np
import torch

import golden_util as gu
from gif_b200.flame_synth import synthetic_flame_model
from oracle import flame_oracle as FO


def _params(g, dt):
    return [torch.from_numpy(g[k]).to(dt) for k in ("shape", "exp", "pose", "eye", "neck")]


def test_oracle_matches_reference_golden_fp32_and_fp64():
    g = gu.load_golden("flame_lbs.npz")
    m = synthetic_flame_model()
    shape, exp, pose, eye, neck = _params(g, torch.float32)
    v, l2, l3 = FO.flame_forward(m, shape, exp, pose, eye, neck)
    assert np.abs(v.numpy() - g["vertices_f32"]).max() < 2e-6
    assert np.abs(l2.numpy() - g["landmarks2d_f32"]).max() < 2e-6
    assert np.abs(l3.numpy() - g["landmarks3d_f32"]).max() < 2e-6
    shape, exp, pose, eye, neck = _params(g, torch.float64)
    v, l2, l3 = FO.flame_forward(m, shape, exp, pose, eye, neck)
    st = int(g["vertices_f64_stride"])
    assert np.abs(v.numpy()[:, ::st] - g["vertices_f64"]).max() < 1e-6      # fp32-stored params, fp64 arithmetic
    assert l2.shape == (4, 68, 3) and l3.shape == (4, 68, 3)


def test_rest_pose_is_the_template_and_rotation_is_rigid():
    m = synthetic_flame_model()
    dt = torch.float64
    z = lambda *s: torch.zeros(*s, dtype=dt)
    v, _, _ = FO.flame_forward(m, z(1, 100), z(1, 50), z(1, 6))
    assert (v[0] - m["v_template"].to(dt)).abs().max() < 1e-7             # (1e-8 epsilon of batch_rodrigues)
    # a global rotation moves every vertex rigidly about the root joint: pairwise distances are preserved
    pose = z(1, 6)
    pose[0, :3] = torch.tensor([0.2, -0.7, 0.1], dtype=dt)
    shape = torch.randn(1, 100, dtype=dt, generator=torch.Generator().manual_seed(1))
    v0, _, _ = FO.flame_forward(m, shape, z(1, 50), z(1, 6))
    v1, _, _ = FO.flame_forward(m, shape, z(1, 50), pose)
    idx = torch.arange(0, v0.shape[1], 97)
    d0 = torch.cdist(v0[0, idx], v0[0, idx])
    d1 = torch.cdist(v1[0, idx], v1[0, idx])
    # the pose blend shapes (posedirs) depend on joints 1.. only, so a pure global rotation is rigid -- up to the 1e-8
    # epsilon batch_rodrigues adds to the (zero) rotation vectors of the other joints
    assert (d0 - d1).abs().max() < 1e-7


def test_synthetic_model_shapes_follow_flame():
    m = synthetic_flame_model()
    V = m["v_template"].shape[0]
    assert V == 5023 and m["faces"].shape == (9976, 3)
    assert m["shapedirs"].shape == (V, 3, 150) and m["posedirs"].shape == (36, V * 3)
    assert m["J_regressor"].shape == (5, V) and m["lbs_weights"].shape == (V, 5)
    assert torch.allclose(m["J_regressor"].sum(1), torch.ones(5), atol=1e-5)
    assert torch.allclose(m["lbs_weights"].sum(1), torch.ones(V), atol=1e-5)
    assert m["parents"].tolist() == [-1, 0, 1, 1, 1]
