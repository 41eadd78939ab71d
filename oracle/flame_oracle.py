"""TEST INFRASTRUCTURE ONLY (see oracle/README / DESIGN.md 6): CPU restatement of the reference's FLAME decoder --
linear blend skinning (my_utils/photometric_optimization/models/lbs.py) and the FLAME.forward wrapper
(my_utils/photometric_optimization/models/FLAME.py:175-216).  Plain torch, any dtype (fp64 arbitrates); pinned by
oracle/make_flame_golden.py against the *unmodified* reference functions (lbs.lbs, lbs.vertices2landmarks,
FLAME._find_dynamic_lmk_idx_and_bcoords) on a synthetic FLAME-shaped model -- the real generic_model.pkl is
licence-gated and absent, so the MODEL is synthetic while the ALGORITHM is the reference's.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module."""
import math

import torch


def batch_rodrigues(rot_vecs):
    """lbs.py:247-279: angle = ||r + 1e-8|| (the epsilon is added to every component), axis = r / angle,
    R = I + sin K + (1 - cos) K K."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    c, s = torch.cos(angle)[:, :, None], torch.sin(angle)[:, :, None]
    z = torch.zeros_like(d[:, 0])
    K = torch.stack([z, -d[:, 2], d[:, 1], d[:, 2], z, -d[:, 0], -d[:, 1], d[:, 0], z], 1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return eye + s * K + (1 - c) * (K @ K)


def batch_rigid_transform(rot_mats, joints, parents):
    """lbs.py:296-349: chain of [R_i | J_i - J_parent] along the kinematic tree; returns posed joints and the transforms
    relative to the rest pose, A_i = [Rc_i | tc_i - Rc_i J_i]."""
    B, N = joints.shape[:2]
    rel = joints.clone()
    rel[:, 1:] = joints[:, 1:] - joints[:, parents[1:]]
    T = torch.zeros(B, N, 4, 4, dtype=joints.dtype)
    T[:, :, :3, :3] = rot_mats
    T[:, :, :3, 3] = rel
    T[:, :, 3, 3] = 1
    chain = [T[:, 0]]
    for i in range(1, N):
        chain.append(chain[int(parents[i])] @ T[:, i])
    chain = torch.stack(chain, 1)
    posed = chain[:, :, :3, 3].clone()
    A = chain.clone()
    A[:, :, :3, 3] = chain[:, :, :3, 3] - (chain[:, :, :3, :3] @ joints[..., None])[..., 0]
    return posed, A


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    """lbs.py:141-228 (pose2rot=True).  betas (B,NB), pose (B,3J) axis-angle, v_template (V,3), shapedirs (V,3,NB),
    posedirs (9(J-1), 3V), J_regressor (J,V), lbs_weights (V,J) -> verts (B,V,3), posed joints (B,J,3)."""
    B = betas.shape[0]
    v_shaped = v_template[None] + torch.einsum("bl,mkl->bmk", betas, shapedirs)          # :176, blend_shapes :232-244
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)                                 # :180, vertices2joints
    R = batch_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)                              # :186-187
    feat = (R[:, 1:] - torch.eye(3, dtype=betas.dtype)).reshape(B, -1)                      # :189
    v_posed = v_shaped + (feat @ posedirs).view(B, -1, 3)                                   # :191-201
    Jt, A = batch_rigid_transform(R, J, parents)                                            # :203
    T = (lbs_weights[None] @ A.view(B, -1, 16)).view(B, -1, 4, 4)                           # :207-211
    homo = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=betas.dtype)], 2)
    verts = (T @ homo[..., None])[:, :, :3, 0]                                              # :213-219
    return verts, Jt


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary_coords):
    """lbs.py:103-138: barycentric interpolation on the listed faces; lmk_faces_idx (B,L), lmk_bary_coords (B,L,3)."""
    tri = faces[lmk_faces_idx]                                         # (B,L,3) vertex ids
    B = vertices.shape[0]
    pts = vertices[torch.arange(B)[:, None, None], tri]               # (B,L,3,3)
    return torch.einsum("blfi,blf->bli", pts, lmk_bary_coords)


def dynamic_landmark_rows(full_pose, neck_kin_chain):
    """FLAME.py:88-132: the contour-landmark table row, from the y rotation (degrees, rounded, clamped at 39) of the
    head relative to the root of the neck chain; negative angles use rows 39..78."""
    B = full_pose.shape[0]
    aa = full_pose.view(B, -1, 3)[:, neck_kin_chain]
    R = batch_rodrigues(aa.reshape(-1, 3)).view(B, -1, 3, 3)
    rel = torch.eye(3, dtype=full_pose.dtype)[None].expand(B, -1, -1)
    for i in range(len(neck_kin_chain)):
        rel = R[:, i] @ rel
    sy = torch.sqrt(rel[:, 0, 0] ** 2 + rel[:, 1, 0] ** 2)                                  # rot_mat_to_euler, FLAME.py:28-34
    ang = torch.round(torch.clamp(torch.atan2(-rel[:, 2, 0], sy) * 180.0 / math.pi, max=39)).long()
    neg = ang < 0
    return torch.where(neg, torch.where(ang < -39, torch.full_like(ang, 78), 39 - ang), ang)


def neck_kin_chain(parents, neck_idx=1):
    """FLAME.py:80-86."""
    chain, cur = [], neck_idx
    while cur != -1:
        chain.append(cur)
        cur = int(parents[cur])
    return torch.tensor(chain, dtype=torch.long)


def flame_forward(model, shape_params, expression_params, pose_params, eye_pose_params=None, neck_pose=None):
    """FLAME.forward (FLAME.py:175-216) on a dict of model tensors (gif_b200.flame_synth.synthetic_flame_model layout):
    betas = [shape | expression]; full pose = [global(3) | neck(3) | jaw(3) | eyes(6)]; -> vertices, landmarks2d (17 dynamic
    contour + 51 static), landmarks3d (68)."""
    dt = shape_params.dtype
    m = {k: (v.to(dt) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in model.items()}
    B = shape_params.shape[0]
    eye = torch.zeros(B, 6, dtype=dt) if eye_pose_params is None else eye_pose_params
    neck = torch.zeros(B, 3, dtype=dt) if neck_pose is None else neck_pose
    betas = torch.cat([shape_params, expression_params], 1)
    full_pose = torch.cat([pose_params[:, :3], neck, pose_params[:, 3:], eye], 1)
    verts, _ = lbs(betas, full_pose, m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"], m["parents"],
                   m["lbs_weights"])
    rows = dynamic_landmark_rows(full_pose, neck_kin_chain(m["parents"]))
    f_idx = torch.cat([m["dynamic_lmk_faces_idx"][rows], m["lmk_faces_idx"][None].expand(B, -1)], 1)
    f_bc = torch.cat([m["dynamic_lmk_bary_coords"][rows], m["lmk_bary_coords"][None].expand(B, -1, -1)], 1)
    lmk2d = vertices2landmarks(verts, m["faces"], f_idx, f_bc)
    lmk3d = vertices2landmarks(verts, m["faces"], m["full_lmk_faces_idx"].expand(B, -1),
                               m["full_lmk_bary_coords"].expand(B, -1, -1))
    return verts, lmk2d, lmk3d
