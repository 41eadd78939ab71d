"""Checkpoint compatibility with the reference: ``train.py:255-262`` writes
    {'generator_running', 'generator', 'g_optimizer', 'discriminator_flm', 'd_optimizer_flm'}
with every network wrapped in ``nn.DataParallel`` (train.py:344,356,358), i.e. ``module.``-prefixed keys, and reads it back at
train.py:389-395; the plot scripts read ``generator_running`` only (plots/generate_random_samples.py:143-144).
The gif_b200 modules keep the reference's state_dict keys, shapes AND parameter order (tests/test_abi.py), so a published
``.model`` file loads with ``strict=True`` once the prefix is stripped, and the Adam states (indexed by parameter position)
carry over as they are.

Order with CUDA graphs: load -> warm-up -> capture is the natural one, but loading into a trainer that HAS captured is
also safe: network weights are copied in place by ``load_state_dict`` and the Adam moments / step counters are copied INTO
the existing state tensors (``torch.optim.Optimizer.load_state_dict`` alone would rebind them to fresh storage while the
captured graphs keep updating the old addresses)."""
import os

import numpy as np
import torch

_NETS = ("generator_running", "generator", "discriminator_flm")


def strip_dataparallel_prefix(state_dict, prefix="module."):
    """``module.x`` -> ``x`` (keys without the prefix are kept: checkpoints of unwrapped models load too)."""
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in state_dict.items()}


def add_dataparallel_prefix(state_dict, prefix="module."):
    return {prefix + k: v for k, v in state_dict.items()}


def load_reference_checkpoint(ckpt, generator=None, g_running=None, discriminator=None, g_optimizer=None, d_optimizer=None,
                              strict=True, map_location="cpu", allow_pickle=False):
    """``ckpt``: a path to a reference ``.model`` file or the dict itself.  Loads whichever objects are given (a plot
    script passes only ``g_running``).  Returns the checkpoint dict.  The format is tensors + Adam state dicts, so the
    safe loader (``weights_only=True``) is the default; ``allow_pickle=True`` opts into full unpickling for a legacy file
    from a trusted source."""
    if not isinstance(ckpt, dict):
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=not allow_pickle)
    missing = [k for k, obj in (("generator", generator), ("generator_running", g_running), ("discriminator_flm", discriminator),
                                ("g_optimizer", g_optimizer), ("d_optimizer_flm", d_optimizer)) if obj is not None and k not in ckpt]
    if missing:
        raise KeyError(f"checkpoint has no {missing}; keys: {sorted(ckpt)}")
    for key, net in (("generator", generator), ("generator_running", g_running), ("discriminator_flm", discriminator)):
        if net is not None:
            net.load_state_dict(strip_dataparallel_prefix(ckpt[key]), strict=strict)
    for key, opt in (("g_optimizer", g_optimizer), ("d_optimizer_flm", d_optimizer)):
        if opt is not None:
            _load_optimizer(opt, ckpt[key])
    return ckpt


def _load_optimizer(opt, state):
    """Adam state by parameter position.  The reference's Adam keeps ``step`` as a host number / CPU tensor; a
    ``capturable=True`` optimiser (the CUDA-graph trainer) needs it on the parameter's device: torch's own
    ``load_state_dict`` does that cast, the hyper-parameters of THIS optimiser (lr, betas, capturable) are kept."""
    state = {"state": state["state"], "param_groups": [dict(g) for g in state["param_groups"]]}
    for g_new, g_old in zip(state["param_groups"], opt.state_dict()["param_groups"]):
        keep = {k: v for k, v in g_old.items() if k != "params"}
        g_new.update(keep)
    live = {p: dict(st) for p, st in opt.state.items()}          # tensors a captured CUDA graph may have baked in
    opt.load_state_dict(state)
    for p, st in opt.state.items():
        for k, v in list(st.items()):
            old = live.get(p, {}).get(k)
            if torch.is_tensor(old) and torch.is_tensor(v) and old.shape == v.shape:
                old.copy_(v)                                      # keep the address, take the value
                st[k] = old


def reference_checkpoint_dict(generator, g_running, discriminator, g_optimizer, d_optimizer, dataparallel_prefix=True):
    """The dict ``train.py:257-261`` saves, from gif_b200 objects (so the reference's own ``--ckpt`` path and plot scripts
    can read what this framework trained)."""
    wrap = add_dataparallel_prefix if dataparallel_prefix else (lambda d: dict(d))
    return {"generator_running": wrap(g_running.state_dict()), "generator": wrap(generator.state_dict()),
            "g_optimizer": g_optimizer.state_dict(), "discriminator_flm": wrap(discriminator.state_dict()),
            "d_optimizer_flm": d_optimizer.state_dict()}


def _sidecar(path):
    return path[:-len(".model")] + ".npz" if path.endswith(".model") else path + ".npz"


def save_trainer(path, trainer, dataparallel_prefix=True, batch=None):
    """Write a GifTrainer's state in the reference's checkpoint format, plus the ``.npz`` sidecar train.py:262-263 writes
    (step, used_sampless, alpha, resolution) extended by ``iteration`` so the every-16th R1 phase survives a resume."""
    torch.save(reference_checkpoint_dict(trainer.generator, trainer.g_running, trainer.discriminator, trainer.g_optimizer,
                                         trainer.d_optimizer, dataparallel_prefix), path)
    np.savez(_sidecar(path), step=trainer.step_idx, used_sampless=trainer.iteration * (batch or 0), alpha=1,
             resolution=4 * 2 ** trainer.step_idx, iteration=trainer.iteration)


def load_trainer(ckpt, trainer, strict=True, allow_pickle=False):
    """Resume a GifTrainer from a reference (or ``save_trainer``) checkpoint; the iteration counter comes from the sidecar
    when there is one (a reference sidecar has no ``iteration``: the counter then restarts at 0, like train.py's ``i``)."""
    out = load_reference_checkpoint(ckpt, trainer.generator, trainer.g_running, trainer.discriminator, trainer.g_optimizer,
                                    trainer.d_optimizer, strict=strict, map_location=trainer.device, allow_pickle=allow_pickle)
    if not isinstance(ckpt, dict) and os.path.isfile(_sidecar(ckpt)):
        side = np.load(_sidecar(ckpt))
        if "iteration" in side.files:
            trainer.iteration = int(side["iteration"])
    return out
