"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules from /root/reference.

Used in the build container to (a) validate the restatement in ``oracle/stylegan2_oracle.py`` and
(b) generate the golden fixtures under ``tests/golden/`` (see ``oracle/make_golden.py``).
``/root/reference`` does not exist on the GPU box, so nothing under ``tests -m gpu``, ``smoke()`` or
``bench.py`` may call this module; it raises if the reference tree is absent.

Import recipe: the reference pulls in visualisation / renderer dependencies at
import time that are irrelevant to the hot path (``my_utils/graph_writer/graph_writer.py:3-5`` needs pyvis +
matplotlib; ``my_utils/photometric_optimization/renderer.py:10-13`` needs pytorch3d + skimage;
``dataset_loaders.py:13`` needs lmdb).  They are replaced by empty stand-ins *in sys.modules only*.
"""
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("GIF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "stylegan2_common_layers.py"))


class _NullSpace:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a namespace with the reference's hot-path modules (cl, gen, disc, losses)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}; ref_import is container-only test tooling")
    # the repo's own `model` shim (gif_b200.install_as_reference_modules) must not shadow the reference
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    gw = _stub("my_utils.graph_writer.graph_writer", ModuleSpace=_NullSpace,
               CallWrapper=lambda obj, node_tracing_name=None: obj, draw=lambda *a, **k: None)
    pkg = _stub("my_utils.graph_writer", graph_writer=gw)
    pkg.__path__ = []
    for n in ("pytorch3d", "pytorch3d.structures", "pytorch3d.io", "pytorch3d.renderer", "pytorch3d.renderer.mesh"):
        _stub(n, Meshes=None, load_obj=None, rasterize_meshes=None).__path__ = []
    _stub("skimage").__path__ = []
    _stub("skimage.io", imread=None, imsave=None)
    _stub("lmdb")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import my_utils  # noqa: F401  (reference package; __path__ stays the reference's)
        from model import stylegan2_common_layers as cl
        from model import stg2_generator as gen
        from model import stg2_discriminator as disc
        from loss_functions import losses
    _loaded.update(cl=cl, gen=gen, disc=disc, losses=losses)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def quiet():
    """The reference constructors print parameter counts (stg2_generator.py:143-155,244)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
