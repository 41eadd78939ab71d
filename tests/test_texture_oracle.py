"""CPU: oracle/texture_oracle.py against the golden written by the unmodified reference
(FlameTextureSpace.compute_texture_map, model/stg2_generator.py:376-421; oracle/make_flame_golden.py)."""
import numpy as np
import torch

import golden_util as gu
from gif_b200.flame_synth import synthetic_flame_model, synthetic_texture_data
from oracle import texture_oracle as TO


def test_texture_oracle_matches_reference_golden():
    g = gu.load_golden("texture_steal.npz")
    td, m = synthetic_texture_data(), synthetic_flame_model()
    verts, cam, src = (torch.from_numpy(g[k]) for k in ("verts", "cam", "src"))
    tex, mask = TO.texture_space_forward(src, verts, cam, m["faces"], td)
    assert np.abs(tex.numpy()[:, :, ::3, ::3] - g["tex_sub"]).max() < 2e-6
    assert np.array_equal(np.packbits(mask.numpy()), g["mask"])
    assert 0.2 < float(mask.float().mean()) < 0.7           # front-facing part of the head


def test_invalid_texels_sample_the_image_centre_and_loss_is_symmetric():
    td, m = synthetic_texture_data(), synthetic_flame_model()
    g = torch.Generator().manual_seed(0)
    src = torch.rand(2, 3, 8, 8, generator=g)
    verts = m["v_template"][None].expand(2, -1, -1).contiguous()
    cam = torch.tensor([[7.0, 0.0, 0.0], [8.0, 0.01, -0.02]])
    tex, mask = TO.texture_space_forward(src, verts, cam, m["faces"], td)
    owner = -np.ones(256 * 256, dtype=np.int64)
    owner[td["y_coords"][td["valid_pixel_ids"]] * 256 + td["x_coords"][td["valid_pixel_ids"]]] = 1
    inv = torch.from_numpy(owner.reshape(256, 256) < 0)
    centre = src[:, :, 3:5, 3:5].mean((2, 3))                # align_corners=False: grid 0 = between the 4 centre pixels
    assert torch.allclose(tex[:, :, inv], centre[:, :, None].expand(-1, -1, int(inv.sum())), atol=1e-6)
    assert not mask[:, 0, inv].any()
    region = torch.ones(1, 1, 256, 256)
    l01 = TO.tex_sp_intrp_loss(tex, mask.float(), [(0, 1)], region)
    l10 = TO.tex_sp_intrp_loss(tex, mask.float(), [(1, 0)], region)
    assert torch.allclose(l01, l10) and float(l01) > 8.0 * 0.999   # sigmoid(0) = 0.5 on every masked-out texel: 16 * 0.5 floor
