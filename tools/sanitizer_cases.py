#!/usr/bin/env python
"""A few small invocations of the round-2 kernels for `compute-sanitizer` (memcheck / racecheck / synccheck): the bf16x3
convolution (S1, S2, T2 incl. the border launches and the split-K schedule + reduction), the bf16x3 weight gradient (plain,
STACK, HALO), the split pass, the fused activation backward with planes, the second-order tail node, the multi-tensor Adam,
the split-K GEMM, and the rasteriser (tile kernel with shared-memory atomics, per-face backward, both
conventions).  Each result is checked against the exact path so a silent out-of-bounds would also show as a mismatch.
usage: compute-sanitizer --tool memcheck python tools/sanitizer_cases.py"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gif_b200 import ops, rasterize  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def conv(B, Hs, Ws, Ci, Co, k, mode):
    Hi, Wi = (Hs, Ws) if mode != 1 else (2 * Hs + 1, 2 * Ws + 1)
    x = torch.randn(B, Hi, Wi, Ci, device=dev, generator=g)
    w = torch.randn(k * k, Co, Ci, device=dev, generator=g) / math.sqrt(Ci * k * k)
    out = []
    for impl in (3, 1):
        ops.CONV_IMPL = impl
        out.append(ops._conv_raw(x, w, k, mode, False, False, (ops.conv_out_size(Hi, k, mode), ops.conv_out_size(Wi, k, mode)))[0])
    e = rel(out[0], out[1])
    assert e < 5e-5, (B, Hs, Ws, Ci, Co, k, mode, e)
    if mode == 0:
        gy = torch.randn_like(out[0])
        res = []
        for impl in (3, 1):
            ops.CONV_IMPL = impl
            res.append(ops._wgrad_raw(x, gy, k, mode, False, False))
        e = rel(res[0], res[1])
        assert e < 5e-5, ("wgrad", B, Hs, Ws, Ci, Co, e)


for case in [(2, 16, 16, 32, 32, 3, 0), (1, 32, 64, 64, 128, 3, 0), (3, 4, 4, 64, 32, 3, 0), (2, 16, 16, 32, 32, 3, 1),
             (2, 8, 32, 32, 32, 3, 2), (2, 16, 16, 128, 128, 3, 0), (1, 8, 8, 96, 128, 1, 0)]:
    conv(*case)
ops.CONV_IMPL = 3
ops._PRECISION = "bf16x3"
# fused activation backward with planes (ConvBiasAct path) through autograd
x = torch.randn(2, 16, 16, 64, device=dev, generator=g).requires_grad_(True)
w = (torch.randn(9, 128, 64, device=dev, generator=g) / 24).requires_grad_(True)
b = torch.zeros(128, device=dev, requires_grad=True)
y = ops.conv2d_bias_act(x, w, b, 3)
y.square().sum().backward()
assert all(torch.isfinite(t.grad).all() for t in (x, w, b))
# second-order tail / modulation node (gifb200_tail_bwd2, vector and scalar variants) as the path-length term drives it
for c in (64, 20):
    acc = torch.randn(2, 9, 7, c, device=dev, generator=g).requires_grad_(True)
    dm = (torch.rand(2, c, device=dev, generator=g) + 1.0).requires_grad_(True)
    nz = torch.randn(2, 9, 7, c, device=dev, generator=g)
    bs = torch.randn(c, device=dev, generator=g)
    gy = torch.randn(2, 9, 7, c, device=dev, generator=g).requires_grad_(True)
    yy = ops.bias_act(acc, bs, 0.2, math.sqrt(2.0), rowscale=dm, add=nz)
    with ops.input_gradient_only():
        ga, gd = torch.autograd.grad(yy, (acc, dm), gy, create_graph=True)
    (ga.square().sum() + gd.square().sum()).backward()
    assert all(torch.isfinite(t.grad).all() for t in (acc, dm, gy))
# multi-tensor Adam (two packs: > 64 tensors, odd sizes, unaligned tails) and the split-K style GEMM (M = 32, K = 512)
from gif_b200.optim import FusedAdam  # noqa: E402
ps = [torch.randn(s, device=dev, generator=g).requires_grad_(True) for s in [(33,), (64, 64, 3, 3), (5, 7), (1,)] + [(17,)] * 70]
opt = FusedAdam(ps, lr=1e-3, betas=(0.0, 0.99))
for _ in range(2):
    for q in ps:
        q.grad = torch.randn(q.shape, device=dev, generator=g)
    opt.step()
assert all(torch.isfinite(q).all() for q in ps)
am, bm = torch.randn(32, 512, device=dev, generator=g), torch.randn(512, 512, device=dev, generator=g)
assert rel(ops.matmul(am, bm, trans_b=True), am @ bm.t()) < 2e-5
ops.set_precision("tf32")
# rasteriser, both conventions, forward + backward
r = np.random.default_rng(0)
fv = np.concatenate([r.uniform(-4, 52, (2, 300, 1, 2)) + r.normal(0, 5, (2, 300, 3, 2)), r.uniform(1, 3, (2, 300, 3, 1))], -1).astype(np.float32)
col = torch.rand(2, 300, 3, 3, device=dev, requires_grad=True)
fvt = torch.from_numpy(fv).to(dev).requires_grad_(True)
d, t, im, im2 = rasterize.rasterize(fvt, 48, 48, col, col * 0.5)
(im.sum() + im2.sum() + d[t >= 0].sum()).backward()
ndc = torch.from_numpy(np.concatenate([r.uniform(-1, 1, (2, 200, 1, 2)) + r.normal(0, 0.1, (2, 200, 3, 2)), r.uniform(0.5, 3, (2, 200, 3, 1))], -1).astype(np.float32)).to(dev).requires_grad_(True)
z, t2, bw = rasterize.rasterize(ndc, 40, 40, convention="pytorch3d")
(bw * (t2 >= 0)[..., None]).sum().backward()
torch.cuda.synchronize()
print("sanitizer cases: OK")
