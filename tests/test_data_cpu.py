"""Host-side input pipeline: the LMDB reader against files produced by the bulk writer (parity unpinned: no lmdb
package and no LMDB file exist in this image; the layout constants are those of lmdb's published format), the reference's key
schema, PIL decode + normalisation, and the pinned batch loader."""
import io
import os
import struct

import numpy as np
import pytest
import torch

from gif_b200 import data


def _png(arr):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(arr).save(b, format="PNG")
    return b.getvalue()


@pytest.mark.parametrize("page_size,n", [(4096, 300), (512, 2500), (1024, 40)])
def test_lmdb_roundtrip_depths_and_overflow(tmp_path, page_size, n):
    rng = np.random.default_rng(n)
    items = {}
    for i in range(n):
        size = int(rng.choice([0, 3, 40, 200, 900, 5000, 70000], p=[.05, .2, .3, .2, .1, .1, .05]))
        items[f"{int(rng.integers(4, 1025))}-{i:05d}".encode()] = rng.integers(0, 256, size, dtype=np.uint8).tobytes()
    items[b"length"] = str(n).encode()
    path = str(tmp_path / "env")
    f = data.write_lmdb(path, items.items(), page_size=page_size)
    # documented layout constants: two meta pages with the magic at byte 16, main-db record at 16 + 72
    raw = open(f, "rb").read()
    assert struct.unpack_from("<I", raw, 16)[0] == 0xBEEFC0DE and struct.unpack_from("<I", raw, page_size + 16)[0] == 0xBEEFC0DE
    assert struct.unpack_from("<H", raw, 10)[0] & 0x08 and len(raw) % page_size == 0
    r = data.LmdbReader(path)
    assert r.page_size == page_size and len(r) == len(items)
    if page_size == 512:
        assert r.meta["depth"] >= 3                                   # branch pages over branch pages
    for k, v in items.items():
        assert r.get(k) == v, k
    assert r.get(b"missing") is None and r.get(b"") is None and r.get(b"\xff" * 20, b"dflt") == b"dflt"
    assert r.get("length") == str(n).encode()
    assert list(r.items()) == sorted(items.items())                  # full scan in key order
    r.close()


def test_lmdb_reader_rejects_garbage(tmp_path):
    p = tmp_path / "data.mdb"
    p.write_bytes(b"\0" * 8192)
    with pytest.raises(data.LmdbFormatError):
        data.LmdbReader(str(p))


def test_dataset_key_schema_decode_and_loader(tmp_path):
    """prepare_ffhq_multiscale_dataset.py:56-61 key schema -> dataset_loaders.py:236-330 item -> batches."""
    n, R = 10, 32
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, (n, R, R, 3), dtype=np.uint8)
    rend = rng.integers(0, 256, (n, 16, 16, 3), dtype=np.uint8)      # stored at a lower resolution: resized on read
    nrm = rng.integers(0, 256, (n, R, R, 3), dtype=np.uint8)
    real_items = [(data.image_key(R, i), _png(imgs[i])) for i in range(n)] + [(b"length", str(n).encode())]
    real_items += [(data.image_key(8, i), _png(imgs[i][::4, ::4])) for i in range(n)]          # multiscale: other sizes coexist
    rend_items = [(data.image_key(16, i), _png(rend[i])) for i in range(n)] + [(data.normal_map_key(16, i), _png(nrm[i][::2, ::2])) for i in range(n)]
    data.write_lmdb(str(tmp_path / "real"), real_items)
    data.write_lmdb(str(tmp_path / "rend"), rend_items)
    assert data.image_key(256, 7) == b"256-00007" and data.normal_map_key(256, 7) == b"norm_map_256-00007"
    flame = rng.normal(size=(n, 159)).astype(np.float32)
    ds = data.GifLmdbDataset(str(tmp_path / "real"), str(tmp_path / "rend"), flame, resolution=R, rend_flm_res=16,
                             flame_mean=0.5, flame_std=2.0)
    assert len(ds) == n
    img, cond, lbl, idx = ds[3]
    assert idx == 3 and tuple(img.shape) == (3, R, R) and tuple(cond[0].shape) == (6, R, R) and tuple(lbl[0].shape) == (159,)
    want = torch.from_numpy(imgs[3].astype(np.float32)).permute(2, 0, 1) / 255.0 * 2 - 1
    assert torch.allclose(img, want, atol=1e-6)                       # ToTensor + Normalize(0.5, 0.5)
    assert torch.allclose(lbl[0], torch.from_numpy((flame[3] - 0.5) / 2.0))
    assert float(cond[0].min()) >= -1 and float(cond[0].max()) <= 1
    loader = data.PinnedBatchLoader(ds, batch_size=4, shuffle=False, pin=False)
    batches = [(a.clone(), b.clone(), c.clone(), d.clone()) for a, b, c, d in loader]
    assert len(batches) == 2                                          # drop_last
    assert batches[1][3].tolist() == [4, 5, 6, 7]
    assert torch.equal(batches[0][0][3], img) and torch.equal(batches[0][1][3], cond[0])
