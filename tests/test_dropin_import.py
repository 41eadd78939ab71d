"""Container / host-side check of the drop-in boundary: with ``gif_b200.install_as_reference_modules()`` active,
the reference's own import chain -- ``from loss_functions import losses`` (losses.py:9 pulls ``FlameTextureSpace`` out of
``model.stg2_generator``) and ``import train`` (train.py:15-21) -- resolves to this repo's modules, and
``nn.DataParallel`` is the one-process-per-GPU stand-in exposing ``.module`` (train.py:344-367).  No kernels are launched.
Runs in a subprocess so that the module table of the test session stays clean.  The GPU counterpart that RUNS the
reference's train() is tests/test_dropin_train_gpu.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oracle import ref_import
assert ref_import.available(), "no reference tree"
train = ref_import.load_train(with_gif_b200=True)
import torch, gif_b200
from gif_b200 import distributed, texture_space
import model.stg2_generator as gen, model.stg2_discriminator as disc, model.stylegan2_common_layers as cl
assert gen.__name__ == "gif_b200.model.stg2_generator" and cl.__name__ == "gif_b200.model.stylegan2_common_layers"
assert train.StyledGenerator is gen.StyledGenerator and train.Discriminator is disc.Discriminator
assert train.stylegan2_common_layers is cl
from loss_functions import losses
assert losses.__file__.startswith(ref_import.REF_ROOT), losses.__file__       # the reference's own losses.py ...
assert losses.FlameTextureSpace is texture_space.FlameTextureSpace              # ... bound to this repo's class
assert train.losses is losses
for name in ("FusedLeakyReLU PixelNorm Upsample Downsample Blur EqualConv2d EqualLinear ScaledLeakyReLU ModulatedConv2d "
             "NoiseInjection ConstantInput StyledConv ToRGB get_w_frm_z ConvLayer ResBlock Generator upfirdn2d make_kernel").split():
    assert hasattr(cl, name), name
for name in "StyledGenerator Generator StyledConvStyleGAN2 ImgEmbedding ConstantInput FlameTextureSpace".split():
    assert hasattr(gen, name), name
assert torch.nn.DataParallel is distributed.DataParallel
net = torch.nn.DataParallel(torch.nn.Linear(3, 2))
assert isinstance(net.module, torch.nn.Linear) and list(net.state_dict()) == ["module.weight", "module.bias"]
assert net(torch.ones(1, 3)).shape == (1, 2)
g = cl.Generator(16, 512, 2)                                                   # the plain StyleGAN2 stack is constructible
assert [k for k in g.state_dict() if k.startswith("noises.")] == ["noises.noise_%%d" %% i for i in range(5)]
print("DROPIN-IMPORT-OK")
"""


def test_reference_import_chain_resolves_to_gif_b200():
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("no reference tree (/root/reference or oracle/_ref/pyref)")
    r = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, os.path.join(ROOT, "tests"))], capture_output=True, text=True,
                       timeout=600)
    assert "DROPIN-IMPORT-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
