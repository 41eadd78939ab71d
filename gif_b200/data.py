"""Input side of the hot path: the reference's on-disk format -> pinned host batches -> the trainer.

The reference keeps its training images and its pre-rendered FLAME conditions in two LMDB environments
(dataset_loaders.py:157-182) written by prepare_lmdb/prepare_ffhq_multiscale_dataset.py:56-61 and
prepare_lmdb/create_deca_rendered_lmdb.py:57-89:

    real images       key  f"{resolution}-{index:05d}"            value  encoded image bytes (PNG / JPEG), RGB
    rendered FLAME    key  f"{resolution}-{index:05d}"            value  encoded texture render
    normal maps       key  f"norm_map_{resolution}-{index:05d}"   value  encoded normal-map render
    "length"          ASCII decimal number of images

and ``FFHQ.__getitem__`` (dataset_loaders.py:236-330) decodes them with PIL, maps to [-1, 1] (``ToTensor`` + ``Normalize(0.5, 0.5)``,
dataset_loaders.py:128-137) and returns ``(img, [cat(render, normal)], [flame_label], index)``.

This module reads that format WITHOUT the ``lmdb`` package (absent from the image; a dependency of the reference, not a
vendored source): ``LmdbReader`` is a read-only walker of LMDB's published file format (data.mdb: two meta pages, a B+tree of
branch / leaf pages with 2-byte node offsets, overflow pages for values larger than half a page), restated from the format
description in lmdb's mdb.c / lmdb.h (MDB_page, MDB_node, MDB_meta, MDB_db).  ``write_lmdb`` is a bulk writer of the same
format used to build fixtures (sorted keys -> packed leaves -> branch levels -> meta pages).  PARITY UNPINNED: no LMDB file and
no lmdb library exist here to check either against; reader and writer are tested against each other and against the
documented layout constants (tests/test_data_cpu.py).

``GifLmdbDataset`` mirrors the reference dataset's item contract for the configuration the flagship run uses (rendered FLAME +
normal maps as condition); ``PinnedBatchLoader`` assembles whole batches in pinned host memory on a background thread, which is
exactly what ``GifTrainer.train_iteration`` / bench.py's ``e2e`` leg consume."""
import io
import os
import queue
import struct
import threading

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------- LMDB file format
PAGE_HDR = 16                      # MDB_page: pgno u64, pad u16, flags u16, (lower u16, upper u16) | overflow page count u32
P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA = 0x01                   # node flag: the value lives in overflow pages, the node holds their first page number
MDB_MAGIC, MDB_VERSION = 0xBEEFC0DE, 1
P_INVALID = 0xFFFFFFFFFFFFFFFF
NODE_HDR = 8                       # MDB_node: lo u16, hi u16, flags u16, ksize u16


class LmdbFormatError(IOError):
    pass


class LmdbReader:
    """Read-only view of an LMDB environment's main database (``path`` = the directory holding data.mdb, or the file itself
    when the environment was created with ``subdir=False``)."""

    def __init__(self, path):
        if os.path.isdir(path):
            path = os.path.join(path, "data.mdb")
        self.path = path
        self._f = open(path, "rb")
        self._buf = np.memmap(path, dtype=np.uint8, mode="r")
        head = bytes(self._buf[:PAGE_HDR + 16])
        if struct.unpack_from("<I", head, PAGE_HDR)[0] != MDB_MAGIC:
            raise LmdbFormatError(f"{path}: not an LMDB data file (bad magic)")
        metas = []
        # the page size is not stored as such: it is the distance between the two meta pages; md_pad of the FREE_DBI record
        # of meta page 0 holds it (mdb.c: mm_psize is an alias of mm_dbs[FREE_DBI].md_pad)
        self.page_size = struct.unpack_from("<I", bytes(self._buf[PAGE_HDR + 24:PAGE_HDR + 28]), 0)[0]
        if self.page_size < 512 or self.page_size & (self.page_size - 1):
            raise LmdbFormatError(f"{path}: implausible page size {self.page_size}")
        for pg in (0, 1):
            off = pg * self.page_size
            flags = struct.unpack_from("<H", bytes(self._buf[off + 10:off + 12]), 0)[0]
            m = bytes(self._buf[off + PAGE_HDR:off + PAGE_HDR + 136])
            magic, version = struct.unpack_from("<II", m, 0)
            if not (flags & P_META) or magic != MDB_MAGIC:
                continue
            if version != MDB_VERSION:
                raise LmdbFormatError(f"{path}: unsupported LMDB data version {version}")
            # mm_address u64, mm_mapsize u64, mm_dbs[2] (48 bytes each), mm_last_pg u64, mm_txnid u64
            main = struct.unpack_from("<IHHQQQQQ", m, 24 + 48)
            last_pg, txnid = struct.unpack_from("<QQ", m, 24 + 96)
            metas.append({"txnid": txnid, "depth": main[2], "entries": main[6], "root": main[7], "last_pg": last_pg})
        if not metas:
            raise LmdbFormatError(f"{path}: no valid meta page")
        self.meta = max(metas, key=lambda d: d["txnid"])

    def close(self):
        self._buf = None
        self._f.close()

    def __len__(self):
        return self.meta["entries"]

    def _page(self, pgno):
        off = pgno * self.page_size
        return memoryview(self._buf[off:off + self.page_size])

    def _nodes(self, page):
        flags, lower = struct.unpack_from("<HH", page, 10)
        n = (lower - PAGE_HDR) // 2
        return flags, struct.unpack_from(f"<{n}H", page, PAGE_HDR)

    @staticmethod
    def _node(page, off):
        lo, hi, nflags, ksize = struct.unpack_from("<HHHH", page, off)
        return lo, hi, nflags, bytes(page[off + NODE_HDR:off + NODE_HDR + ksize])

    def _value(self, page, off, lo, hi, nflags, ksize):
        size = lo | (hi << 16)
        start = off + NODE_HDR + ksize
        if nflags & F_BIGDATA:
            pgno = struct.unpack_from("<Q", page, start)[0]
            o = pgno * self.page_size
            ov_flags, = struct.unpack_from("<H", bytes(self._buf[o + 10:o + 12]), 0)
            if not ov_flags & P_OVERFLOW:
                raise LmdbFormatError(f"page {pgno} is not an overflow page")
            return bytes(self._buf[o + PAGE_HDR:o + PAGE_HDR + size])
        return bytes(page[start:start + size])

    def get(self, key, default=None):
        """The value stored under ``key`` (bytes), or ``default``.  Keys compare as byte strings (LMDB's default order)."""
        if isinstance(key, str):
            key = key.encode("utf-8")
        pgno = self.meta["root"]
        if pgno == P_INVALID:
            return default
        while True:
            page = self._page(pgno)
            flags, ptrs = self._nodes(page)
            if flags & P_BRANCH:
                # child i covers keys >= separator i (separator 0 is empty): the last separator <= key
                lo_i, hi_i = 0, len(ptrs) - 1
                while lo_i < hi_i:
                    mid = (lo_i + hi_i + 1) // 2
                    if self._node(page, ptrs[mid])[3] <= key:
                        lo_i = mid
                    else:
                        hi_i = mid - 1
                lo, hi, nflags, _ = self._node(page, ptrs[lo_i])
                pgno = lo | (hi << 16) | (nflags << 32)
            elif flags & P_LEAF:
                lo_i, hi_i = 0, len(ptrs) - 1
                while lo_i <= hi_i:
                    mid = (lo_i + hi_i) // 2
                    lo, hi, nflags, k = self._node(page, ptrs[mid])
                    if k == key:
                        return self._value(page, ptrs[mid], lo, hi, nflags, len(k))
                    if k < key:
                        lo_i = mid + 1
                    else:
                        hi_i = mid - 1
                return default
            else:
                raise LmdbFormatError(f"page {pgno}: unexpected flags {flags:#x}")

    def items(self):
        """All (key, value) pairs in key order."""
        def walk(pgno):
            page = self._page(pgno)
            flags, ptrs = self._nodes(page)
            for off in ptrs:
                lo, hi, nflags, k = self._node(page, off)
                if flags & P_BRANCH:
                    yield from walk(lo | (hi << 16) | (nflags << 32))
                else:
                    yield k, self._value(page, off, lo, hi, nflags, len(k))
        if self.meta["root"] != P_INVALID:
            yield from walk(self.meta["root"])


def write_lmdb(path, items, page_size=4096):
    """Bulk-write ``items`` (an iterable of (key bytes, value bytes)) as an LMDB environment directory ``path`` (data.mdb).
    Packed leaves, branch levels built bottom-up, values that do not fit half a page go to overflow pages -- the layout a
    reader of the published format (and ``LmdbReader``) expects.  Fixture / export tool, not a transactional store."""
    items = sorted((k if isinstance(k, bytes) else k.encode("utf-8"), bytes(v)) for k, v in items)
    for (a, _), (b, _) in zip(items, items[1:]):
        if a == b:
            raise ValueError(f"duplicate key {a!r}")
    os.makedirs(path, exist_ok=True)
    node_max = ((page_size - PAGE_HDR) // 2 - 2) & ~1            # mdb.c: me_nodemax
    pages = {}                                                     # pgno -> bytes
    next_pg = [2]

    def alloc(n=1):
        p = next_pg[0]
        next_pg[0] += n
        return p

    def even(n):
        return (n + 1) & ~1

    def build_level(entries, leaf):
        """entries: leaf -> (key, node payload bytes, node flags, data size); branch -> (key, child pgno).  Returns
        [(first key, pgno)] of the pages written."""
        out, cur, used = [], [], 0
        cap = page_size - PAGE_HDR

        def flush():
            nonlocal cur, used
            if not cur:
                return
            pgno = alloc()
            page = bytearray(page_size)
            upper = page_size
            ptrs = []
            for idx, e in enumerate(cur):
                if leaf:
                    key, payload, nflags, dsize = e
                    lo, hi = dsize & 0xFFFF, dsize >> 16
                else:
                    key, child = e
                    if idx == 0:
                        key = b""                                  # a branch page's first separator is implicit
                    payload, lo, hi, nflags = b"", child & 0xFFFF, (child >> 16) & 0xFFFF, child >> 32
                node = struct.pack("<HHHH", lo, hi, nflags, len(key)) + key + payload
                upper -= even(len(node))
                page[upper:upper + len(node)] = node
                ptrs.append(upper)
            lower = PAGE_HDR + 2 * len(ptrs)
            assert lower <= upper
            struct.pack_into("<QHHHH", page, 0, pgno, 0, P_LEAF if leaf else P_BRANCH, lower, upper)
            struct.pack_into(f"<{len(ptrs)}H", page, PAGE_HDR, *ptrs)
            pages[pgno] = bytes(page)
            out.append((cur[0][0], pgno))
            cur, used = [], 0
        for e in entries:
            ksz = len(e[0])
            nsz = even(NODE_HDR + ksz + (len(e[1]) if leaf else 0)) + 2
            if used + nsz > cap:
                flush()
            cur.append(e)
            used += nsz
        flush()
        return out
    leaf_entries, overflow_pages = [], 0
    for k, v in items:
        if len(k) > 511:
            raise ValueError("LMDB keys are at most 511 bytes")
        if NODE_HDR + len(k) + len(v) > node_max:
            n = (PAGE_HDR + len(v) + page_size - 1) // page_size
            pg = alloc(n)
            blob = bytearray(n * page_size)
            struct.pack_into("<QHHI", blob, 0, pg, 0, P_OVERFLOW, n)
            blob[PAGE_HDR:PAGE_HDR + len(v)] = v
            for j in range(n):
                pages[pg + j] = bytes(blob[j * page_size:(j + 1) * page_size])
            overflow_pages += n
            leaf_entries.append((k, struct.pack("<Q", pg), F_BIGDATA, len(v)))
        else:
            leaf_entries.append((k, v, 0, len(v)))
    level = build_level(leaf_entries, True)
    leaf_pages, branch_pages, depth = len(level), 0, 1 if level else 0
    while len(level) > 1:
        level = build_level(level, False)
        branch_pages += len(level)
        depth += 1
    root = level[0][1] if level else P_INVALID
    last_pg = next_pg[0] - 1
    with open(os.path.join(path, "data.mdb"), "wb") as f:
        for pg in (0, 1):
            page = bytearray(page_size)
            struct.pack_into("<QHHHH", page, 0, pg, 0, P_META, 0, 0)
            free_db = struct.pack("<IHHQQQQQ", page_size, 0, 0, 0, 0, 0, 0, P_INVALID)
            main_db = struct.pack("<IHHQQQQQ", 0, 0, depth, branch_pages, leaf_pages, overflow_pages, len(items), root)
            meta = struct.pack("<IIQQ", MDB_MAGIC, MDB_VERSION, 0, max(1 << 20, (last_pg + 1) * page_size)) + free_db + main_db + \
                struct.pack("<QQ", last_pg, pg)                    # txnid: page 1 is the newer one
            page[PAGE_HDR:PAGE_HDR + len(meta)] = meta
            f.write(page)
        for pg in range(2, last_pg + 1):
            f.write(pages[pg])
    return os.path.join(path, "data.mdb")


# ---------------------------------------------------------------------------------------------- key schema / decode
def image_key(resolution, index):
    """prepare_ffhq_multiscale_dataset.py:58, dataset_loaders.py:252."""
    return f"{resolution}-{str(index).zfill(5)}".encode("utf-8")


def normal_map_key(resolution, index):
    """dataset_loaders.py:262."""
    return f"norm_map_{resolution}-{str(index).zfill(5)}".encode("utf-8")


def decode_image(data, resolution=None):
    """Encoded bytes -> float32 (3,H,W) in [-1,1]: PIL decode, optional resize (dataset_loaders.py:268-270,276-279),
    ``ToTensor`` + ``Normalize((0.5,)*3, (0.5,)*3)`` (dataset_loaders.py:128-137)."""
    from PIL import Image
    img = Image.open(io.BytesIO(data)).convert("RGB")
    if resolution is not None and img.size[0] != resolution:
        img = img.resize((resolution, resolution))
    a = np.asarray(img, dtype=np.float32)
    return torch.from_numpy(a).permute(2, 0, 1).div_(255.0).sub_(0.5).div_(0.5)


class GifLmdbDataset(torch.utils.data.Dataset):
    """The reference's FFHQ item for ``rendered_flame_as_condition=True, normal_maps_as_cond=True`` (dataset_loaders.py:236-330):
    ``(img (3,R,R), [cond (6,R,R)], [flame_label (P,)], index)``, all float32, images in [-1,1]."""

    def __init__(self, real_img_root, rendered_flame_root, flame_params, resolution=256, rend_flm_res=256, valid_ids=None,
                 flame_mean=0.0, flame_std=1.0):
        self.real = LmdbReader(real_img_root)
        self.rend = LmdbReader(rendered_flame_root)
        self.length = int(self.real.get(b"length").decode("utf-8"))           # dataset_loaders.py:168
        self.resolution, self.rend_flm_res = resolution, rend_flm_res
        self.flame_params = np.asarray(flame_params, dtype=np.float32)
        self.valid_ids = np.arange(self.length) if valid_ids is None else np.asarray(valid_ids)
        self.flame_mean, self.flame_std = flame_mean, flame_std

    def __len__(self):
        return len(self.valid_ids)

    def __getitem__(self, index):
        i = int(self.valid_ids[index])
        img = decode_image(self.real.get(image_key(self.resolution, i)))
        rnd = decode_image(self.rend.get(image_key(self.rend_flm_res, i)), self.resolution)
        nrm = decode_image(self.rend.get(normal_map_key(self.rend_flm_res, i)), self.resolution)
        lbl = (self.flame_params[i] - self.flame_mean) / self.flame_std
        return img, [torch.cat((rnd, nrm), 0)], [torch.from_numpy(np.asarray(lbl, dtype=np.float32))], i


class PinnedBatchLoader:
    """Batches of a GifLmdbDataset assembled in PINNED host memory by a background thread (decode is host work; the copy to the
    device is one non-blocking transfer per tensor, issued by the trainer).  Yields (real (B,3,R,R), cond (B,6,R,R),
    labels (B,P), indices (B,) int64) -- the arguments of ``GifTrainer.train_iteration``; ``depth`` batches are in flight."""

    def __init__(self, dataset, batch_size, shuffle=True, seed=0, depth=3, pin=None):
        self.ds, self.bs, self.shuffle, self.seed, self.depth = dataset, batch_size, shuffle, seed, depth
        self.pin = torch.cuda.is_available() if pin is None else pin
        img, cond, lbl, _ = dataset[0]
        self._shapes = (tuple(img.shape), tuple(cond[0].shape), tuple(lbl[0].shape))

    def _alloc(self):
        mk = (lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype).pin_memory()) if self.pin else \
            (lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype))
        return (mk(self.bs, *self._shapes[0]), mk(self.bs, *self._shapes[1]), mk(self.bs, *self._shapes[2]),
                mk(self.bs, dtype=torch.int64))

    def __iter__(self):
        order = np.arange(len(self.ds))
        if self.shuffle:
            np.random.default_rng(self.seed).shuffle(order)
            self.seed += 1
        nb = len(order) // self.bs                                              # drop_last=True, dataset_loaders.py:395
        free, ready = queue.Queue(), queue.Queue(maxsize=self.depth)
        for _ in range(self.depth + 1):
            free.put(self._alloc())

        def work():
            for b in range(nb):
                buf = free.get()
                for j, idx in enumerate(order[b * self.bs:(b + 1) * self.bs]):
                    img, cond, lbl, i = self.ds[int(idx)]
                    buf[0][j].copy_(img); buf[1][j].copy_(cond[0]); buf[2][j].copy_(lbl[0]); buf[3][j] = i
                ready.put(buf)
            ready.put(None)
        threading.Thread(target=work, daemon=True).start()
        while True:
            buf = ready.get()
            if buf is None:
                return
            yield buf
            free.put(buf)       # the consumer is done with the previous batch once it asks for the next one
