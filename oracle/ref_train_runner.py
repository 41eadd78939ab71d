"""TEST INFRASTRUCTURE ONLY -- runs the reference's UNMODIFIED ``train()`` (train.py:28-312) on synthetic batches.

Consumers: tests/dropin_harness.py (loop-body parity and drop-in boundary tests), oracle/make_train_golden.py, and
bench.py's CPU legs (``cpu_baseline`` / ``--impl reference``: the reference's own modules and loop timed on the host
cores).  Never the product path.

``train()`` reads four globals that train.py defines under ``__main__`` (g_optimizer, d_optimizer_flm, g_running, n_critic;
train.py:322,365-382) and three names it imports (``sample_data``, ``tqdm``): the runner sets them on the imported module
exactly as ``__main__`` / the imports would, with a synthetic loader and a silent progress bar.  The progress-bar stand-in
also chooses where the loop counter starts (``first_i``): train.py:79 iterates ``tqdm(range(3_000_000))``, so handing the
loop ``range(15, ...)`` makes the FIRST iteration an R1 iteration (``(i + 1) % 16 == 0``, train.py:145)."""
import contextlib
import math
import time
import types

import torch


class StopTraining(Exception):
    pass


class SyntheticDataset:
    """What train() touches of the dataset object (train.py:120-122, :230, :269)."""

    def __init__(self, on_iteration_start=None):
        self.calls = 0
        self.on_iteration_start = on_iteration_start

    def accumulate_batches_of_flm(self, flm, pose):     # called once per iteration, after the batch was drawn
        if self.on_iteration_start is not None:
            self.on_iteration_start(self.calls)          # == number of COMPLETED iterations
        self.calls += 1

    def un_normalize_flame(self, x):
        return x


def make_args(res, batch, vocab):
    return types.SimpleNamespace(
        embedding_vocab_size=vocab, gen_reg_type="None", batch={res: batch}, batch_default=batch, debug=True,
        lr={}, use_styled_conv_stylegan2=True, max_size=res, init_size=res, phase=10 ** 9, ckpt=None,
        rendered_flame_as_condition=True, normal_maps_as_cond=True, shfld_cond_as_neg_smpl=False, embedding_reg_weight=0.0,
        apply_texture_space_interpolation_loss=False, adaptive_interp_loss=False, use_posed_constant_input=False, run_id="t")


@contextlib.contextmanager
def cuda_calls_are_noops_without_a_gpu(force=False):
    """train() calls ``.cuda()`` on its batches (train.py:125-130).  Without a GPU (the build container) -- or with
    ``force`` (the CPU legs of bench.py, which time the reference on the host cores of a GPU box) -- those become no-ops so
    that the unmodified function runs on the CPU."""
    if torch.cuda.is_available() and not force:
        yield
        return
    t_cuda, m_cuda, avail = torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.is_available
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    # nn.DataParallel moves its module to cuda:0 with ``.to()`` when exactly one GPU is visible (and scatters the inputs
    # there): with a GPU in the box the "CPU" arm silently ran on it through cuDNN.  No visible device type -> device_ids = []
    # -> DataParallel.forward calls the module where it is.
    torch.cuda.is_available = lambda: False
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.is_available = t_cuda, m_cuda, avail


def run(train_mod, G, D, Gr, batches, res, vocab, first_i=0, on_iteration_done=None, force_cpu=False, stamps=None):
    """Runs train_mod.train() over ``batches`` (a list of (real, cond, labels, indices) host tensors, one per iteration).
    ``on_iteration_done(k, generator, discriminator, g_running, g_optimizer, d_optimizer)`` is called after k = 1, 2, ...
    completed iterations.  ``stamps``: a list that receives time.perf_counter() at the start of every iteration and at the
    end of the last one (so ``stamps[k+1] - stamps[k]`` is the wall time of iteration k)."""
    with cuda_calls_are_noops_without_a_gpu(force_cpu):
        return _run(train_mod, G, D, Gr, batches, res, vocab, first_i, on_iteration_done, stamps)


def _run(train_mod, G, D, Gr, batches, res, vocab, first_i, on_iteration_done, stamps):
    from torch import nn, optim
    n_iters = len(batches)
    batch_size = batches[0][0].shape[0]
    generator = nn.DataParallel(G).cuda()                  # train.py:344 (the drop-in's stand-in when it is installed)
    discriminator = nn.DataParallel(D).cuda()              # train.py:356
    g_running = nn.DataParallel(Gr).cuda()                 # train.py:358
    g_running.train(False)
    g_ratio, d_ratio = 4 / 5, 16 / 17                      # train.py:365-382
    train_mod.g_optimizer = optim.Adam(generator.module.parameters(), lr=0.002 * g_ratio, betas=(0.0, 0.99 ** g_ratio))
    train_mod.d_optimizer_flm = optim.Adam(discriminator.parameters(), lr=0.002 * d_ratio, betas=(0.0, 0.99 ** d_ratio))
    train_mod.g_running = g_running
    train_mod.n_critic = 1                                 # train.py:322

    def sample_data(dataset, bs, image_sizes, debug=False):
        assert bs == batch_size and image_sizes[-1] == res

        class Loader:
            def __iter__(self):
                def gen():
                    for real, cond, lbls, idx in batches:
                        if stamps is not None:
                            stamps.append(time.perf_counter())
                        yield real, [cond], [lbls], idx
                    if stamps is not None:
                        stamps.append(time.perf_counter())
                    raise StopTraining()
                return gen()
        return Loader()

    class Bar:                                             # ``pbar = tqdm(range(...))`` then ``pbar.set_description``
        def __init__(self, it):
            self.it = it

        def __iter__(self):
            return iter(range(first_i, len(self.it)))

        def set_description(self, *_a, **_k):
            pass

    train_mod.sample_data = sample_data
    train_mod.tqdm = Bar

    def done(k):
        if k > 0 and on_iteration_done is not None:
            on_iteration_done(k, generator.module, discriminator.module, g_running.module, train_mod.g_optimizer,
                              train_mod.d_optimizer_flm)

    dataset = SyntheticDataset(on_iteration_start=done)
    try:
        train_mod.train(make_args(res, batch_size, vocab), dataset, generator, discriminator, None, None, 0,
                        int(math.log2(res)) - 2)
    except StopTraining:
        pass
    done(n_iters)
    return generator.module, discriminator.module, g_running.module
