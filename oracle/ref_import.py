"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules from /root/reference.

Used in the build container to (a) validate the restatement in ``oracle/stylegan2_oracle.py`` and
(b) generate the golden fixtures under ``tests/golden/`` (see ``oracle/make_golden.py``).
``/root/reference`` does not exist on the GPU box; there the module falls back to ``oracle/_ref/pyref`` -- the files of
this import chain copied verbatim at build time by ``oracle/extract_pyref.py`` (git-ignored, never committed, exactly
like ``oracle/_ref/ref_kernels.inc``) -- and raises if neither is present.  Consumers: tests, ``bench.py --impl reference``
and the ``cpu_baseline`` leg (the reference timed on the host cores).  Never the product path.

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

_HERE = os.path.dirname(os.path.abspath(__file__))
PYREF = os.path.join(_HERE, "_ref", "pyref")     # build-time extract of the reference's python files (oracle/extract_pyref.py)
REF_ROOT = os.environ.get("GIF_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REF_ROOT, "model", "stylegan2_common_layers.py")) and \
        os.path.isfile(os.path.join(PYREF, "model", "stylegan2_common_layers.py")):
    REF_ROOT = PYREF        # the GPU box: /root/reference does not exist there, the git-ignored extract travels with gpurun


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


def _install_stubs():
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


class _NoViz:
    """Stand-in for my_utils.visualize_flame_overlay.OverLayViz in the train() harness: its constructor loads the
    licence-gated FLAME model (visualize_flame_overlay.py:13-16); train()'s loop body never calls it (only the sample
    saver every 500th iteration does)."""

    def __init__(self, *a, **k):
        pass


def load_train(with_gif_b200):
    """Imports the reference's UNMODIFIED train.py (train.py:1-403) as a module and returns it.

    with_gif_b200=True : ``gif_b200.install_as_reference_modules()`` is active, so train.py's ``from model... import`` and
                         ``nn.DataParallel`` resolve to this repo's drop-in modules -- the boundary test.
    with_gif_b200=False: the reference's own model/ and loss modules (CPU oracle of the loop body, train.py:80-252).
    Stubs (sys.modules only): graph_writer / pytorch3d / skimage / lmdb as in ``load``; ``imageio``; ``my_utils.compute_fid``
    (FID runs every 500th iteration, outside the loop body); ``OverLayViz`` (see _NoViz).  ``train()`` reads the globals
    ``g_optimizer``, ``d_optimizer_flm``, ``g_running``, ``n_critic`` that train.py defines under ``__main__``
    (train.py:322-382): the caller sets them on the returned module, as ``__main__`` would."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    purge = ("train", "loss_functions", "my_utils.generic_utils", "my_utils.visualize_flame_overlay", "dataset_loaders")
    for k in [k for k in sys.modules if k in purge or k.startswith("loss_functions.")]:
        del sys.modules[k]
    if not with_gif_b200:
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
    _install_stubs()
    _stub("imageio")
    _stub("my_utils.compute_fid", FidComputer=None)
    if with_gif_b200:
        import gif_b200
        gif_b200.install_as_reference_modules()
    with contextlib.redirect_stdout(io.StringIO()):
        import my_utils  # noqa: F401
        import my_utils.visualize_flame_overlay as vfo
        vfo.OverLayViz = _NoViz
        import train
    return train


def load():
    """Returns a namespace with the reference's hot-path modules (cl, gen, disc, losses)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}; ref_import is container-only test tooling")
    # the repo's own `model` shim (gif_b200.install_as_reference_modules) must not shadow the reference
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    _install_stubs()
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
