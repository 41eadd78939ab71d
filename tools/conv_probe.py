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
         "s1": (ops.S1, 32, 128, 256, 256), "s2small": (ops.S2, 32, 9, 512, 512), "t2small": (ops.T2, 32, 8, 512, 512)}


def main(names):
    ops.set_precision("tf32")
    for n in names:
        mode, b, r, ci, co = CASES[n]
        x = ops._round_tf32_raw(torch.randn(b, r, r, ci, device=dev))
        w = torch.randn(9, co, ci, device=dev) / math.sqrt(9 * ci)
        ho = ops.conv_out_size(r, 3, mode)
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
