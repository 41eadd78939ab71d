"""CPU: the rasteriser C oracle against (a) the goldens produced by the UNMODIFIED reference kernels executed on the CPU
and (b), when oracle/_ref is present (build container, or shipped prebuilt to the GPU box), those kernels live;
and against the reference's only golden artefact, body_vis.obj."""
import numpy as np
import pytest

import golden_util as gu
import raster_cases
from oracle import rasterize_oracle as RO


@pytest.mark.parametrize("name", list(raster_cases.all_cases().keys()))
def test_oracle_matches_reference_kernels(name):
    fv, colors, h, w = raster_cases.all_cases()[name]
    g = gu.load_golden("raster_cases.npz")
    d, t, o = (RO.oracle_rasterize_colors(fv, colors, h, w) if colors is not None else RO.oracle_rasterize(fv, h, w))
    assert np.array_equal(d, g[name + "_depth"])
    assert np.array_equal(t, g[name + "_tri"])
    assert np.array_equal(o, g[name + "_out3"])
    if RO.have_ref():
        dr, tr, orr = (RO.ref_rasterize_colors(fv, colors, h, w) if colors is not None else RO.ref_rasterize(fv, h, w))
        assert np.array_equal(dr, d)
        tie = tr != t
        assert np.array_equal(tie, g[name + "_tie"])       # index differences are exactly the exact-zp ties
        assert np.array_equal(orr[~tie], o[~tie])


def test_body_visibility():
    g = gu.load_golden("body_visibility.npz")
    vis = RO.get_visibility(g["vertices"][None], g["faces"][None].astype(np.int64), 512, 512)
    assert int((vis[0] != g["visible"]).sum()) == 0
    assert int(g["visible"].sum()) == 3723
