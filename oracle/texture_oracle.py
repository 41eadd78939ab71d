"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's texture "stealing" (FlameTextureSpace.compute_texture_map,
model/stg2_generator.py:376-421, and .forward :355-374) and of InterpolatedTextureLoss.pairwise_texture_loss /
tex_sp_intrp_loss (loss_functions/losses.py:147-176).  Pinned by oracle/make_flame_golden.py against the unmodified
reference method on the synthetic texture table (gif_b200.flame_synth.synthetic_texture_data).
Only tests/, smoke() and bench.py's cpu_baseline may import this module."""
import torch
import torch.nn.functional as F


def batch_orth_proj(X, camera):
    """model/mesh_and_3d_helpers.py:40-50: scale * (x + tx, y + ty, z)."""
    camera = camera.reshape(-1, 1, 3)
    return camera[:, :, 0:1] * torch.cat([X[:, :, :2] + camera[:, :, 1:], X[:, :, 2:]], 2)


def vertex_normals(vertices, faces):
    """model/mesh_and_3d_helpers.py:5-37 (faces (F,3) shared by the batch)."""
    B, V = vertices.shape[:2]
    vf = vertices[:, faces]
    n = torch.zeros(B, V, 3, dtype=vertices.dtype)
    n.index_add_(1, faces[:, 1], torch.cross(vf[:, :, 2] - vf[:, :, 1], vf[:, :, 0] - vf[:, :, 1], dim=-1))
    n.index_add_(1, faces[:, 2], torch.cross(vf[:, :, 0] - vf[:, :, 2], vf[:, :, 1] - vf[:, :, 2], dim=-1))
    n.index_add_(1, faces[:, 0], torch.cross(vf[:, :, 1] - vf[:, :, 0], vf[:, :, 2] - vf[:, :, 0], dim=-1))
    return F.normalize(n, eps=1e-6, dim=2)


def compute_texture_map(source_img, verts, vnormals, cam, td, size=256):
    """stg2_generator.py:376-421.  source_img (B,C,H,W); verts / vnormals (B,V,3); cam (B,3) = (scale, tx, ty);
    td: the texture table (numpy arrays).  Texels without a triangle keep grid (0,0) = the image centre (as written);
    the visibility mask is `normal_z < 0` on valid texels, False elsewhere."""
    vid = torch.as_tensor(td["valid_pixel_3d_faces"]).long()
    bc = torch.as_tensor(td["valid_pixel_b_coords"]).to(verts.dtype)
    ys = torch.as_tensor(td["y_coords"][td["valid_pixel_ids"]]).long()
    xs = torch.as_tensor(td["x_coords"][td["valid_pixel_ids"]]).long()
    p3d = sum(verts[:, vid[:, k], :] * bc[:, k][None, :, None] for k in range(3))
    proj = batch_orth_proj(p3d, cam)[:, :, :2].clone()
    proj[:, :, 1] *= -1
    B = source_img.shape[0]
    grid = torch.zeros(B, size, size, 2, dtype=source_img.dtype)
    grid[:, ys, xs, :] = proj.to(source_img.dtype)
    tex = F.grid_sample(source_img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    pn = sum(vnormals[:, vid[:, k], :] * bc[:, k][:, None] for k in range(3))
    mask = torch.zeros(B, 1, size, size, dtype=torch.bool)
    mask[:, :, ys, xs] = (pn[:, :, -1:] < 0).transpose(1, 2)
    return tex, mask


def texture_space_forward(source_img, verts, cam, faces, td):
    """FlameTextureSpace.forward after the FLAME decode (stg2_generator.py:366-374): normals of the projected, y/z-flipped mesh."""
    tv = batch_orth_proj(verts, cam).clone()
    tv[:, :, 1:] = -tv[:, :, 1:]
    return compute_texture_map(source_img, verts, vertex_normals(tv, faces), cam, td)


def pairwise_texture_loss(tx1, tx2, region_mask):
    """losses.py:147-159: mean(sigmoid((tx1 - tx2)^2) * mask); region_mask (1,1,H,W) -> its [0] broadcasts over the batch-less
    (C,H,W) textures."""
    return torch.mean(torch.sigmoid(torch.pow(tx1 - tx2, 2)) * region_mask[0])


def tex_sp_intrp_loss(textures, tx_masks, pairs, region_mask):
    """losses.py:166-176 for a given list of index pairs: 16 * sum_pairs L(tx_i * m, tx_j * m) / len(pairs), m = m_i * m_j."""
    loss = 0
    for i, j in pairs:
        m = tx_masks[j] * tx_masks[i]
        loss = loss + pairwise_texture_loss(textures[i] * m, textures[j] * m, region_mask)
    return 16 * loss / len(pairs)
