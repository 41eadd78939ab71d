"""TEST INFRASTRUCTURE -- drives the reference's UNMODIFIED ``train()`` (train.py:28-312) on a synthetic dataset.

Used three ways, always with the same seeded inputs and initial weights:
  * ``oracle/make_train_golden.py`` (build container): train() over the reference's OWN model / loss modules on the CPU
    -> ``tests/golden/train_loop.npz`` (parameters, EMA generator, Adam moments after 1 and after 16 iterations; the 16th
    is the first R1 iteration, train.py:145);
  * ``tests/test_dropin_train_gpu.py``: the same train() over THIS repo's drop-in modules
    (``gif_b200.install_as_reference_modules()``: model.* and nn.DataParallel resolve here) -> must reproduce the golden:
    the boundary test of SURVEY 8b and the loop-body parity of 8a L4 in one;
  * the same test file runs ``gif_b200.train_step.GifTrainer`` (one shared G forward, flat gradient buffers, fused Adam)
    on the same batches against the same golden.
The runner that drives train() (globals it reads, synthetic loader, loop-counter start) is oracle/ref_train_runner.py.
"""
import contextlib

import numpy as np
import torch

import golden_util as gu

RES, BATCH, VOCAB, ITERS = 16, 4, 50, 16
G_SEED, D_SEED, RUNNING_SEED = 11, 12, 13

G_WATCH = ["generator.progression.0.st_cv1.conv.weight", "generator.progression.2.st_cv2.conv.weight",
           "generator.progression.1.st_cv1.conv.modulation.weight", "generator.progression.2.st_cv1.noise.noise_conv.4.weight",
           "generator.progression.2.st_cv2.activate.bias", "generator.to_rgb.2.conv.weight", "generator.to_rgb.1.bias",
           "generator.const_input.input", "z_to_w.1.weight", "z_to_w.8.bias"]
D_WATCH = ["convs.0.0.weight", "convs.0.1.bias", "convs.1.conv1.0.weight", "convs.1.conv2.1.weight", "convs.2.skip.1.weight",
           "convs.2.conv2.2.bias", "final_conv.0.weight", "final_linear.0.weight", "final_linear.1.weight", "final_linear.1.bias"]


def batch(i, device="cpu"):
    """Iteration i's batch: FFHQ-shaped image in [-1,1], 6-channel condition in [-1,1], FLAME labels, identity indices."""
    real = gu.rand_uniform((BATCH, 3, RES, RES), 7000 + 4 * i)
    cond = gu.rand_uniform((BATCH, 6, RES, RES), 7001 + 4 * i)
    lbls = gu.randn((BATCH, 159), 7002 + 4 * i)
    idx = gu.randint(VOCAB, (BATCH,), 7003 + 4 * i)
    return real.to(device), cond.to(device), lbls.to(device), idx.to(device)


def initial_state_dicts():
    g = gu.seeded_state_dict(gu.g_shapes(VOCAB), G_SEED)
    d = gu.seeded_state_dict(gu.d_shapes(RES), D_SEED)
    r = gu.seeded_state_dict(gu.g_shapes(VOCAB), RUNNING_SEED)
    r["image_embedding.embd_weight"] = r["img_embdng.embd_weight"] = g["image_embedding.embd_weight"]
    return g, d, r


def build_networks(gen_mod, disc_mod, device):
    """StyledGenerator x2 + Discriminator from the given (reference or drop-in) modules, seeded weights."""
    g_sd, d_sd, r_sd = initial_state_dicts()
    kw = dict(embedding_vocab_size=VOCAB, rendered_flame_ascondition=True, normal_maps_as_cond=True, core_tensor_res=4, n_mlp=8)
    with contextlib.redirect_stdout(None):
        G, Gr = gen_mod.StyledGenerator(**kw), gen_mod.StyledGenerator(**kw)
        D = disc_mod.Discriminator(size=RES, num_color_chnls=9, channel_multiplier=2)
    G.load_state_dict(g_sd)
    Gr.load_state_dict(r_sd)
    D.load_state_dict(d_sd)
    return G.to(device), D.to(device), Gr.to(device)


def run_reference_train(train_mod, G, D, Gr, n_iters, snapshot_after=(1,), first_i=0):
    """Runs train_mod.train() for n_iters iterations on batch(first_i), batch(first_i + 1), ...  Returns {k: snapshot} for
    k in snapshot_after + (n_iters,).  ``first_i``: where train()'s loop counter starts (oracle/ref_train_runner.py): with
    first_i = 15 the FIRST iteration is an R1 iteration (train.py:145) of the unmodified function."""
    from oracle import ref_train_runner
    snaps = {}

    def on_done(k, g, d, gr, g_opt, d_opt):
        if (k in snapshot_after or k == n_iters) and k not in snaps:
            snaps[k] = snapshot(g, d, gr, g_opt, d_opt)

    ref_train_runner.run(train_mod, G, D, Gr, [batch(i) for i in range(first_i, first_i + n_iters)], RES, VOCAB,
                         first_i=first_i, on_iteration_done=on_done)
    return snaps


def snapshot(G, D, Gr, g_opt, d_opt):
    """Sampled parameters / EMA parameters / Adam second moments as float64 numpy (+ norms)."""
    out = {}
    for tag, net, names in (("g", G, G_WATCH), ("d", D, D_WATCH), ("r", Gr, G_WATCH)):
        named = dict(net.named_parameters())
        for n in names:
            s, tot = gu.sample(named[n], 1024, 5)
            out[f"{tag}|{n}"] = s
            out[f"{tag}|{n}|norm"] = np.array(float(named[n].detach().double().norm()))
        flat = torch.cat([p.detach().reshape(-1).double().cpu() for p in net.parameters()])
        out[f"{tag}|all"] = gu.sample(flat, 16384, 6)[0]
        out[f"{tag}|all|norm"] = np.array(float(flat.norm()))
    for tag, opt, net, names in (("g", g_opt, G, G_WATCH), ("d", d_opt, D, D_WATCH)):
        named = dict(net.named_parameters())
        for n in names:
            st = opt.state.get(named[n])
            if st:
                out[f"{tag}|{n}|exp_avg_sq"] = gu.sample(st["exp_avg_sq"], 1024, 7)[0]
    return out


def flatten_snaps(snaps):
    return {f"it{k}|{n}": v for k, s in snaps.items() for n, v in s.items()}


def compare(snap, golden, it, tol_d, tol_g, tol_moment, floor_factor=3.0, scenario=""):
    """L2-relative deviation of every watched tensor from the golden.  The bar of a tensor is max(the net's bar,
    floor_factor x the reference's OWN fp32-vs-fp64 deviation for that tensor at that iteration) -- the golden stores the
    float64 run of the same unmodified train() beside the float32 one, so the chaotic part of a 16-step Adam trajectory
    (tiny tensors such as a 3-element to_rgb bias flip whole +-lr steps) is measured, not guessed."""
    worst = {}
    for key, got in snap.items():
        if key.endswith("|norm"):
            continue
        ref = golden[f"{scenario}it{it}|{key}"]
        ref64 = golden[f"f64|{scenario}it{it}|{key}"]
        floor = float(np.linalg.norm(ref - ref64) / max(np.linalg.norm(ref64), 1e-300))
        tag = key.split("|")[0]
        base = (tol_moment or 0.0) if key.endswith("exp_avg_sq") else (tol_d if tag == "d" else tol_g)
        err = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
        worst[key] = (err, max(base, floor_factor * floor))
    bad = {k: v for k, v in worst.items() if not v[0] < v[1]}
    return worst, bad
