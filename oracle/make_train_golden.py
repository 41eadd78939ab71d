#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY (build container) -- golden for the loop-body parity (train.py:80-252).

Runs the reference's UNMODIFIED ``train()`` over the reference's OWN modules on the CPU (fp32, as the reference trains) for
16 iterations of the synthetic dataset of ``tests/dropin_harness.py`` (16^2, batch 4; the 16th iteration carries the R1
penalty, train.py:145) and stores sampled parameters of G / D / the EMA generator and Adam second moments after iteration 1
and after iteration 16 -> tests/golden/train_loop.npz; plus a second scenario in which the loop counter starts at 15, i.e.
the very first iteration is an R1 iteration (``r1|it1|...``), for a tight check of the penalty's double backward through the
whole loop body.  A second run in float64 is stored beside it (``f64|...``) to show
how far fp32 training trajectories drift by themselves over 16 Adam steps (the tolerance floor of the GPU test)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import dropin_harness as H  # noqa: E402
from oracle import ref_import  # noqa: E402


def run(dtype, n_iters=H.ITERS, first_i=0):
    """Weights and batches are always DRAWN in fp32 (same numbers in both runs), then cast."""
    train = ref_import.load_train(with_gif_b200=False)
    ref = ref_import.load()
    G, D, Gr = H.build_networks(ref.gen, ref.disc, "cpu")
    G, D, Gr = G.to(dtype), D.to(dtype), Gr.to(dtype)
    orig = H.batch
    if dtype == torch.float64:
        H.batch = lambda i, device="cpu": tuple(t.double() if t.is_floating_point() else t for t in orig(i, device))
    try:
        snaps = H.run_reference_train(train, G, D, Gr, n_iters, snapshot_after=(1,), first_i=first_i)
    finally:
        H.batch = orig
    return H.flatten_snaps(snaps)


def main():
    torch.manual_seed(0)
    g32 = run(torch.float32)
    g64 = run(torch.float64)
    # second scenario: the loop counter starts at 15, so the ONE iteration run is an R1 iteration (train.py:145) and uses
    # batch(15); stored under "r1|..." (snapshot after that iteration = "r1|it1|...")
    r32 = {"r1|" + k: v for k, v in run(torch.float32, n_iters=1, first_i=15).items()}
    r64 = {"r1|" + k: v for k, v in run(torch.float64, n_iters=1, first_i=15).items()}
    g32.update(r32)
    g64.update(r64)
    out = dict(g32)
    drift = {}
    for k, v in g64.items():
        out["f64|" + k] = v
        if not k.endswith("|norm"):
            drift[k] = float(np.linalg.norm(g32[k] - v) / max(np.linalg.norm(v), 1e-300))
    ks = [k for k in drift if k.startswith("r1|it1|")]
    print(f"R1-first scenario, fp32-vs-fp64 reference deviation after the iteration: max {max(drift[k] for k in ks):.3e}")
    for it in (1, H.ITERS):
        for tag in "gdr":
            ks = [k for k in drift if k.startswith(f"it{it}|{tag}|") and not k.endswith("exp_avg_sq")]
            print(f"fp32-vs-fp64 reference drift after iteration {it}, net {tag}: max L2-rel over watched tensors "
                  f"{max(drift[k] for k in ks):.3e}")
            out[f"drift|it{it}|{tag}"] = np.array(max(drift[k] for k in ks))
        ks = [k for k in drift if k.startswith(f"it{it}|") and k.endswith("exp_avg_sq")]
        print(f"   Adam exp_avg_sq drift after iteration {it}: {max(drift[k] for k in ks):.3e}")
        out[f"drift|it{it}|moment"] = np.array(max(drift[k] for k in ks))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train_loop.npz"), **out)
    print("wrote tests/golden/train_loop.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
