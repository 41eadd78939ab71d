#!/usr/bin/env python
"""Single weight-gradient launches through the C ABI for ncu / timing: python tools/wgrad_probe.py [ns|mid|low|s2|t2] ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gif_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
CASES = {"ns": (ops.S1, 32, 256, 128, 128), "mid": (ops.S1, 32, 128, 256, 256), "low": (ops.S1, 32, 64, 512, 512),
         "s2": (ops.S2, 32, 257, 128, 256), "t2": (ops.T2, 32, 128, 256, 128)}


def main(names):
    ops.set_precision("tf32")
    for n in names:
        mode, b, r, ci, co = CASES[n]
        ho = ops.conv_out_size(r, 3, mode)
        x = ops._round_tf32_raw(torch.randn(b, r, r, ci, device=dev))
        gy = ops._round_tf32_raw(torch.randn(b, ho, ho, co, device=dev))
        fn = lambda: ops._wgrad_raw(x, gy, 3, mode, False, False)
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
        sites = b * (r * r if mode == ops.T2 else ho * ho)
        print(f"{n}: {ms:.3f} ms  {2.0 * sites * ci * co * 9 / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
