#!/usr/bin/env python
"""Guidance only (numbers taken under a profiler are never bench values): per-kernel device-time table of ONE eager
training iteration (normal / with R1 / with PPL) from torch.profiler (CUPTI), to see where a step goes.
usage: python tools/torch_profile_step.py [normal|r1|ppl] > gpurun_out/step_kernels.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from gif_b200 import ops  # noqa: E402
from gif_b200.train_step import GifTrainer  # noqa: E402


def main(kind):
    dev = torch.device("cuda:0")
    ops.set_precision(os.environ.get("GIFB200_PROBE_PRECISION", "bf16x3"))
    B, R = 32, 256
    tr = GifTrainer(dev, R, ppl=(kind == "ppl"))
    g = torch.Generator(device=dev).manual_seed(0)
    real = torch.rand(B, 3, R, R, device=dev, generator=g) * 2 - 1
    cond = torch.rand(B, 6, R, R, device=dev, generator=g) * 2 - 1
    idx = torch.randint(0, 70000, (B,), device=dev, generator=g)
    for _ in range(3):
        tr.iteration = 0
        tr.train_iteration(real, cond, idx)
    torch.cuda.synchronize()
    tr.iteration = 15 if kind == "r1" else 0
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.train_iteration(real, cond, idx)
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.device_time_total for e in ev)
    print(f"# {kind} iteration [{ops.get_precision()}]: {sum(e.count for e in ev)} kernels, {tot / 1e3:.2f} ms device time")
    for e in sorted(ev, key=lambda e: -e.device_time_total)[:60]:
        print(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / tot:5.1f}%  n={e.count:4d}  avg={e.device_time_total / e.count:8.1f} us  {e.key[:90]}")


if __name__ == "__main__":
    for k in (sys.argv[1:] or ["normal"]):
        main(k)
