"""Builds gif_b200/libgifb200.so (hand-written CUDA for sm_100a + the C ABI of include/gifb200.h) with nvcc.

In-tree build: the .so lands next to this file so that it travels to the GPU box with the gpurun snapshot.
``python -m gif_b200.build`` or ``__graft_entry__.build()``.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libgifb200.so")
SOURCES = ["elementwise.cu", "upfirdn2d.cu", "sgemm.cu", "conv_simt.cu", "conv_tc.cu", "conv_wgrad_tc.cu", "conv_api.cu", "rasterize.cu", "render.cu", "flame.cu", "texture.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _stamp(path):
    h = hashlib.sha1()
    for f in [path, os.path.join(ROOT, "include", "gifb200.h")] + \
            [os.path.join(CSRC, x) for x in sorted(os.listdir(CSRC)) if x.endswith(".cuh")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    objs, rebuilt = [], False
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        stamp_file = obj + ".stamp"
        stamp = _stamp(path)
        if force or not os.path.isfile(obj) or not os.path.isfile(stamp_file) or open(stamp_file).read() != stamp:
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                print(r.stderr)
            with open(stamp_file, "w") as f:
                f.write(stamp)
            rebuilt = True
        objs.append(obj)
    if rebuilt or not os.path.isfile(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
