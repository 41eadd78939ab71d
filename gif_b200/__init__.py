"""gif_b200 -- B200-native (sm_100a) implementation of GIF's data-parallel hot path.

* ``gif_b200.model.stylegan2_common_layers`` / ``stg2_generator`` / ``stg2_discriminator``: drop-in operator and
  model classes (reference API, reference state_dict keys).
* ``gif_b200.losses``: R1 gradient penalty, path-length regulariser, logistic losses (loss_functions/losses.py).
* ``gif_b200.rasterize``: ``standard_rasterize`` / ``standard_rasterize_colors`` (+ differentiable wrapper).
* ``gif_b200.distributed``: one-process-per-GPU gradient all-reduce (replaces train.py's nn.DataParallel).
* ``install_as_reference_modules()``: makes ``from model import ...`` / ``from model.stg2_generator import ...`` in
  the reference's train.py resolve to this package.
"""
import sys

from . import _lib  # noqa: F401  -- loads (or builds) libgifb200.so; import fails loudly without it
from . import ops  # noqa: F401

__version__ = "0.1.0"


def install_as_reference_modules():
    """Register this package's modules under the reference's import names (``model``, ``model.stg2_generator``...)."""
    from . import model
    from .model import stg2_discriminator, stg2_generator, stylegan2_common_layers
    sys.modules["model"] = model
    sys.modules["model.stylegan2_common_layers"] = stylegan2_common_layers
    sys.modules["model.stg2_generator"] = stg2_generator
    sys.modules["model.stg2_discriminator"] = stg2_discriminator
    return model
