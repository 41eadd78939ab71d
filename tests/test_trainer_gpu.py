"""GPU: the training-iteration driver (gif_b200/train_step.py): eager vs CUDA-graph replay give the same losses and
parameters, the R1 / PPL variants run, parameters and the EMA move, unused-resolution parameters get zero gradients."""
import copy

import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu


def _batch(seed, b, res, dev):
    return (gu.rand_uniform((b, 3, res, res), seed).to(dev), gu.rand_uniform((b, 6, res, res), seed + 1).to(dev),
            gu.randint(16, (b,), seed + 2).to(dev))


@pytest.mark.parametrize("ppl", [False, True])
def test_eager_and_graph_iterations_agree(cuda, ppl):
    """d_loss of an iteration depends only on the weights at its start, so it is compared tightly; g_loss is evaluated
    after D's Adam step (beta1 = 0: a normalised-sign update that amplifies summation-order noise of tiny gradients),
    so it is compared loosely, and not at all with PPL (its noise image comes from a different RNG stream in a replay)."""
    from gif_b200 import ops
    from gif_b200.train_step import GifTrainer
    ops.set_precision("tf32")
    res, b = 32, 4
    t_e = GifTrainer(cuda, res, vocab=16, r1_every=2, ppl=ppl, seed=3)
    t_g = GifTrainer(cuda, res, vocab=16, r1_every=2, ppl=ppl, seed=3)
    p0 = copy.deepcopy(t_e.generator.state_dict())

    def sync_weights():
        t_e.generator.load_state_dict(t_g.generator.state_dict())
        t_e.discriminator.load_state_dict(t_g.discriminator.state_dict())

    def both(it):
        sync_weights()
        oe = [float(v) for v in t_e.train_iteration(*_batch(10 * it, b, res, cuda))]
        og = [float(v) for v in t_g.train_iteration(*_batch(10 * it, b, res, cuda))]
        assert all(v == v and abs(v) < 1e4 for v in oe + og), (oe, og)
        assert oe[0] == pytest.approx(og[0], rel=2e-4, abs=1e-5), (it, oe, og)
        if not ppl:
            assert oe[1] == pytest.approx(og[1], rel=5e-2, abs=1e-3), (it, oe, og)

    for it in range(3):          # eager vs eager (iterations 2 of each pair carry R1): the two trainers are the same program
        both(it)
    t_g.capture(b, res)
    assert t_g._graphs is not None and t_g.graph_launches[True] > t_g.graph_launches[False] > 100
    for it in range(3, 7):       # eager vs CUDA-graph replay, both R1 variants, fresh inputs every iteration
        both(it)
    torch.cuda.synchronize()
    # parameters moved, the EMA moved, parameters of unused resolutions (progression.4+) untouched
    sd = t_g.generator.state_dict()
    k_used, k_unused = "generator.progression.2.st_cv1.conv.weight", "generator.progression.5.st_cv1.conv.weight"
    assert not torch.equal(sd[k_used], p0[k_used])
    assert torch.equal(sd[k_unused], p0[k_unused])
    assert not torch.equal(t_g.g_running.state_dict()[k_used], p0[k_used])
    assert all(torch.isfinite(p).all() for p in t_g.discriminator.parameters())
    assert all(torch.isfinite(p).all() for p in t_g.generator.parameters())


def test_texture_interpolation_loss_in_the_iteration(cuda):
    """texture_loss=batch: the G step adds InterpolatedTextureLoss (train.py:224-238) on interpolated FLAME labels -- FLAME
    decode, condition render, a second generator forward, texture stealing, pairwise loss -- eagerly and from CUDA graphs."""
    from gif_b200 import ops
    from gif_b200.train_step import GifTrainer
    ops.set_precision("tf32")
    res, b = 32, 4
    tr = GifTrainer(cuda, res, vocab=16, r1_every=2, ppl=False, seed=5, texture_loss=b)
    ref = GifTrainer(cuda, res, vocab=16, r1_every=2, ppl=False, seed=5)
    g = torch.Generator().manual_seed(2)
    flm = torch.cat([torch.randn(b, 150, generator=g), (torch.rand(b, 6, generator=g) * 2 - 1) * 0.3,
                     torch.rand(b, 1, generator=g) * 3 + 7, (torch.rand(b, 2, generator=g) * 2 - 1) * 0.02], 1).to(cuda)
    outs = []
    for it in range(2):
        real, cond, idx = _batch(10 * it, b, res, cuda)
        d0, g0 = (float(v) for v in ref.train_iteration(real, cond, idx))
        d1, g1 = (float(v) for v in tr.train_iteration(real, cond, idx, flm))
        outs.append((d0, g0, d1, g1))
    # iteration 0: identical weights -> identical D loss; the texture term is positive (>= 16 * 0.5 * masked-out fraction)
    assert outs[0][0] == pytest.approx(outs[0][2], rel=2e-4)
    assert outs[0][3] > outs[0][1] + 1.0
    tr.capture(b, res)
    for it in range(2, 5):
        d1, g1 = (float(v) for v in tr.train_iteration(*_batch(10 * it, b, res, cuda), flm))
        assert d1 == d1 and g1 == g1 and 0 < g1 < 1e3
    assert all(torch.isfinite(p).all() for p in tr.generator.parameters())
