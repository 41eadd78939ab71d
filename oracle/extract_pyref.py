#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- build-time extract of the reference's python files on the hot path's import chain.

``/root/reference`` exists only in the build container.  The parity tests that drive the reference's UNMODIFIED ``train()``
through this repo's drop-in modules (tests/test_dropin_train*.py) and the CPU arm of ``bench.py`` (``--impl reference``: the
reference's own model/ + loss modules timed on the host cores) must also run on the GPU box, so -- exactly like
``oracle/Makefile`` does for the rasteriser kernel text (``oracle/_ref/ref_kernels.inc``) -- this recipe copies the files
VERBATIM into ``oracle/_ref/pyref/`` (git-ignored: never committed, but not gpurun-ignored, so it travels).

Which files: whatever the import of the harness (``oracle/ref_import.load`` + ``load_train``) actually pulls from the
reference tree (discovered from ``sys.modules``), nothing else.  Run by ``__graft_entry__.build()`` when /root/reference is
present; a no-op otherwise (the prebuilt extract is kept)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GIF_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref", "pyref")


def main():
    if not os.path.isfile(os.path.join(REF, "train.py")):
        print("reference tree absent: keeping prebuilt oracle/_ref/pyref (if any)")
        return 0
    sys.path.insert(0, ROOT)
    os.environ["GIF_REFERENCE_ROOT"] = REF
    from oracle import ref_import
    ref_import.load()
    ref_import.load_train(with_gif_b200=False)
    files = set()
    for m in list(sys.modules.values()):
        f = getattr(m, "__file__", None)
        if f and os.path.abspath(f).startswith(os.path.abspath(REF) + os.sep):
            files.add(os.path.abspath(f))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for f in sorted(files):
        rel = os.path.relpath(f, REF)
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copyfile(f, os.path.join(DST, rel))
    print(f"extracted {len(files)} reference python files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
