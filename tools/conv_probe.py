#!/usr/bin/env python
"""Runs a few single convolutions through the C ABI for ncu captures / timing: python tools/conv_probe.py [t2|s2|s1] ..."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gif_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
CASES = {"t2": (ops.T2, 32, 64, 512, 256), "t2b": (ops.T2, 32, 128, 256, 128), "s2": (ops.S2, 32, 129, 256, 512),
         "s1": (ops.S1, 32, 128, 256, 256), "s2small": (ops.S2, 32, 9, 512, 512), "t2small": (ops.T2, 32, 8, 512, 512),
         "northstar": (ops.S1, 32, 256, 128, 128)}


def main(names):
    """GIFB200_PROBE_PRECISION = tf32 (default) | bf16x3; a name with the suffix ":wgrad" times the weight gradient."""
    prec = os.environ.get("GIFB200_PROBE_PRECISION", "tf32")
    ops.set_precision(prec)
    for n in names:
        wgrad = n.endswith(":wgrad")
        mode, b, r, ci, co = CASES[n.split(":")[0]]
        x = torch.randn(b, r, r, ci, device=dev)
        if prec == "tf32":
            x = ops._round_tf32_raw(x)
        w = torch.randn(9, co, ci, device=dev) / math.sqrt(9 * ci)
        ho = ops.conv_out_size(r, 3, mode)
        if wgrad:
            gy = torch.randn(b, ho, ho, co, device=dev)
            fn = lambda: ops._wgrad_raw(x, gy, 3, mode, False, False)
        else:
            fn = lambda: ops._conv_raw(x, w, 3, mode, False, False, (ho, ho))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        taps = 9
        sites = b * (r * r if mode == ops.T2 else ho * ho)
        fl = 2.0 * sites * ci * co * taps
        print(f"{n}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
