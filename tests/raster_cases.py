"""Seeded mesh generators for the rasteriser parity tests (shared by oracle/make_raster_golden.py and tests/).

Face vertices are in PIXEL space as the reference kernel expects (kernel.cu:112-136): x in [0,w], y in [0,h],
z > 0, pixel centres at integer coordinates.
"""
import numpy as np


def random_soup(batch, ntri, h, w, seed, size=6.0, with_colors=False):
    """Random small triangles (both windings -> about half are back faces)."""
    r = np.random.default_rng(seed)
    c = r.uniform([-4, -4], [w + 4, h + 4], size=(batch, ntri, 1, 2))
    xy = c + r.normal(0, size, size=(batch, ntri, 3, 2))
    z = r.uniform(1.0, 3.0, size=(batch, ntri, 3, 1))
    fv = np.concatenate([xy, z], -1).astype(np.float32)
    colors = r.uniform(0, 1, size=(batch, ntri, 3, 3)).astype(np.float32) if with_colors else None
    return fv, colors


def grid_mesh(batch, n, h, w, seed, jitter=0.0, integer=True):
    """A regular n x n grid of quads split into 2 triangles, vertices ON pixel centres when integer=True: pixel
    centres fall exactly on shared edges / vertices, which exercises the asymmetric inside test
    (bw2>=0, bw1>=0, bw0>0, kernel.cu:144) and exact-zp ties between neighbours."""
    r = np.random.default_rng(seed)
    xs = np.linspace(2, w - 3, n + 1)
    ys = np.linspace(2, h - 3, n + 1)
    if integer:
        xs, ys = np.round(xs), np.round(ys)
    gx, gy = np.meshgrid(xs, ys)
    fvs = []
    for _ in range(batch):
        px = gx + r.normal(0, jitter, gx.shape)
        py = gy + r.normal(0, jitter, gy.shape)
        pz = 2.0 + 0.5 * np.sin(gx / 7.0) * np.cos(gy / 5.0) + r.uniform(0, 0.01, gx.shape)
        v = np.stack([px, py, pz], -1)
        tris = []
        for i in range(n):
            for j in range(n):
                a, b, c, d = v[i, j], v[i, j + 1], v[i + 1, j], v[i + 1, j + 1]
                tris.append([a, c, b])      # winding chosen so that the triangles are front-facing (:33)
                tris.append([b, c, d])
        fvs.append(np.asarray(tris))
    return np.asarray(fvs, np.float32), None


def degenerate(h, w):
    """Zero-area, collinear, off-screen, huge, and negative-coordinate triangles."""
    t = [
        [[5, 5, 1], [5, 5, 1], [5, 5, 1]],                      # point
        [[1, 1, 1], [5, 5, 1], [9, 9, 1]],                      # collinear (den == 0 -> inverDeno = 0, :96-99)
        [[-50, -50, 2], [-50, 300, 2], [300, -50, 2]],          # covers the whole image, one winding ...
        [[-50, -50, 3], [300, -50, 3], [-50, 300, 3]],          # ... and the other
        [[w + 5, 2, 1], [w + 9, 2, 1], [w + 5, 9, 1]],          # fully off-screen
        [[3.5, 3.5, 1.5], [3.5, 12.5, 1.5], [12.5, 3.5, 1.5]],  # small, half-integer corners
        [[3.5, 3.5, 1.2], [12.5, 3.5, 1.2], [3.5, 12.5, 1.2]],
        [[20, 20, 1], [20, 30, 1], [30, 20, 1]],                # integer corners (edges through pixel centres)
        [[20, 20, 1], [30, 20, 1], [20, 30, 1]],
    ]
    return np.asarray([t], np.float32), None


def all_cases():
    c = {}
    c["soup_b2_64"] = random_soup(2, 400, 64, 64, 1) + (64, 64)
    c["soup_colors_b3_48x80"] = random_soup(3, 300, 48, 80, 2, with_colors=True) + (48, 80)
    c["soup_big_b1_128"] = random_soup(1, 200, 128, 128, 3, size=30.0) + (128, 128)
    c["grid_integer_b2_64"] = grid_mesh(2, 12, 64, 64, 4) + (64, 64)
    c["grid_jitter_b2_96"] = grid_mesh(2, 20, 96, 96, 5, jitter=0.7, integer=False) + (96, 96)
    c["degenerate_33x40"] = degenerate(33, 40) + (33, 40)
    fv, col = random_soup(1, 1, 16, 16, 6)
    c["empty_zero_faces"] = (fv[:, :0], None, 16, 16)
    return c
