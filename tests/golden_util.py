"""Shared helpers for the parity tests and for ``oracle/make_golden.py``.

* ``g_shapes`` / ``d_shapes``: the reference's ``state_dict`` key -> shape manifest for ``StyledGenerator`` /
  ``Discriminator`` (checked against the unmodified reference in ``oracle/make_golden.py`` and against this
  repo's modules in ``tests/test_state_dict.py``).
* ``seeded_state_dict``: deterministic weights from a CPU ``torch.Generator`` (identical in the build
  container and on the GPU box: same image, same torch), with the reference's initialisation *scales* so the
  networks are numerically well-conditioned.
* ``sample``: a fixed pseudo-random subsample of a tensor (goldens store samples + a float64 sum, not MBs).
"""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

G_CHANNELS = [512, 512, 512, 512, 512, 256, 128, 64, 32]   # progression i -> out channels (gen.py:84-113)
D_CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}


def g_shapes(vocab=100, n_mlp=8, cond_ch=6):
    s = {"generator.const_input.input": (1, 512, 4, 4)}

    def styled(prefix, ci, co, up):
        s[prefix + "conv.weight"] = (1, co, ci, 3, 3)
        if up:
            s[prefix + "conv.blur.kernel"] = (4, 4)
        s[prefix + "conv.modulation.weight"] = (ci, 512)
        s[prefix + "conv.modulation.bias"] = (ci,)
        s[prefix + "noise.noise_conv.0.weight"] = (2 * cond_ch, cond_ch, 3, 3)
        s[prefix + "noise.noise_conv.0.bias"] = (2 * cond_ch,)
        s[prefix + "noise.noise_conv.2.weight"] = (4 * cond_ch, 2 * cond_ch, 3, 3)
        s[prefix + "noise.noise_conv.2.bias"] = (4 * cond_ch,)
        s[prefix + "noise.noise_conv.4.weight"] = (co, 4 * cond_ch, 3, 3)
        s[prefix + "noise.noise_conv.4.bias"] = (co,)
        s[prefix + "activate.bias"] = (1, co, 1, 1)

    ci = 512
    for i, co in enumerate(G_CHANNELS):
        p = f"generator.progression.{i}."
        if i == 0:
            styled(p + "st_cv1.", 512, co, False)
        else:
            styled(p + "st_cv1.", ci, co, True)
            styled(p + "st_cv2.", co, co, False)
        ci = co
    for i, c in enumerate(G_CHANNELS):
        p = f"generator.to_rgb.{i}."
        s[p + "bias"] = (1, 3, 1, 1)
        if i > 0:
            s[p + "upsample.kernel"] = (4, 4)
        s[p + "conv.weight"] = (1, 3, c, 1, 1)
        s[p + "conv.modulation.weight"] = (c, 512)
        s[p + "conv.modulation.bias"] = (c,)
    s["image_embedding.embd_weight"] = (vocab, 512)
    s["img_embdng.embd_weight"] = (vocab, 512)          # same buffer registered twice (gen.py:229-231)
    for i in range(1, n_mlp + 1):
        s[f"z_to_w.{i}.weight"] = (512, 512)
        s[f"z_to_w.{i}.bias"] = (512,)
    return s


def d_shapes(size=256, in_ch=9):
    s = {"convs.0.0.weight": (D_CHANNELS[size], in_ch, 1, 1), "convs.0.1.bias": (1, D_CHANNELS[size], 1, 1)}
    ci = D_CHANNELS[size]
    j = 1
    r = size
    while r > 4:
        co = D_CHANNELS[r // 2]
        p = f"convs.{j}."
        s[p + "conv1.0.weight"] = (ci, ci, 3, 3)
        s[p + "conv1.1.bias"] = (1, ci, 1, 1)
        s[p + "conv2.0.kernel"] = (4, 4)
        s[p + "conv2.1.weight"] = (co, ci, 3, 3)
        s[p + "conv2.2.bias"] = (1, co, 1, 1)
        s[p + "skip.0.kernel"] = (4, 4)
        s[p + "skip.1.weight"] = (co, ci, 1, 1)
        ci = co
        j += 1
        r //= 2
    s["final_conv.0.weight"] = (512, ci + 1, 3, 3)
    s["final_conv.1.bias"] = (1, 512, 1, 1)
    s["final_linear.0.weight"] = (512, 512 * 16)
    s["final_linear.0.bias"] = (512,)
    s["final_linear.1.weight"] = (1, 512)
    s["final_linear.1.bias"] = (1,)
    return s


def blur_kernel(gain=1.0):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = torch.outer(k, k)
    return k / k.sum() * gain


def seeded_state_dict(shapes, seed, dtype=torch.float32):
    """Deterministic weights with the reference's init scales (cl.py:198,296,390-394; gen.py:237)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in shapes:                      # insertion order of the manifest is the draw order
        shp = shapes[k]
        if k == "img_embdng.embd_weight":
            v = sd["image_embedding.embd_weight"]
        elif k.endswith("blur.kernel") or k.endswith("upsample.kernel"):
            v = blur_kernel(4.0)
        elif k.endswith(".kernel"):
            v = blur_kernel(1.0)
        else:
            v = torch.randn(*shp, generator=g)
            if "noise_conv" in k:
                v = v * (0.01 if k.endswith("weight") else 0.05)
            elif k.startswith("z_to_w") and k.endswith("weight"):
                v = v * 100.0
            elif k.endswith("modulation.bias"):
                v = 1.0 + 0.1 * v
            elif k.endswith("bias"):
                v = 0.1 * v
        sd[k] = v.to(dtype)
    return sd


def sample(t, n=2048, seed=0):
    """Fixed pseudo-random subsample (as float64 numpy) of a tensor + its float64 sum."""
    f = t.detach().reshape(-1).to(torch.float64).cpu()
    g = torch.Generator().manual_seed(1000 + seed)
    idx = torch.randint(0, f.numel(), (min(n, f.numel()),), generator=g)
    return f[idx].numpy(), float(f.sum())


def sample_like(t, n=2048, seed=0):
    return sample(t, n, seed)[0]


def rel_err(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the 1e-3 bar of BASELINE.json is read this way)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / den


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def randn(shape, seed, dtype=torch.float32):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).to(dtype)


def rand_uniform(shape, seed, lo=-1.0, hi=1.0, dtype=torch.float32):
    return (torch.rand(*shape, generator=torch.Generator().manual_seed(seed)) * (hi - lo) + lo).to(dtype)


def randint(high, shape, seed):
    return torch.randint(0, high, shape, generator=torch.Generator().manual_seed(seed))
