"""GPU: gif_b200.optim.FusedAdam (multi-tensor gifb200_adam_step) against torch.optim.Adam, the optimiser the reference
constructs (train.py:365-382): same trajectory, same state layout, state_dict round trips in both directions, CUDA-graph
replays advance the device-side step counter."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(512, 512, 3, 3), (1, 512, 1, 1), (3,), (7, 5), (129,), (1,), (256, 3, 3, 3), (1000, 512), (2, 3, 5, 7)] + [(17,)] * 70


def _params(dev, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in SHAPES]


def _set_grads(ps, seed):
    g = torch.Generator().manual_seed(seed)
    for p in ps:
        p.grad = (torch.randn(p.shape, generator=g) * (10.0 ** float(torch.randint(-4, 2, (1,), generator=g)))).to(p.device)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("betas", [(0.0, 0.99 ** (4 / 5)), (0.9, 0.999)])
def test_fused_adam_follows_torch_adam(cuda, betas):
    from gif_b200.optim import FusedAdam
    ours, ref = _params(cuda, 1), _params(cuda, 1)
    o = FusedAdam(ours, lr=0.002 * 4 / 5, betas=betas)
    r = torch.optim.Adam(ref, lr=0.002 * 4 / 5, betas=betas, foreach=False)
    for it in range(6):
        _set_grads(ours, 100 + it)
        _set_grads(ref, 100 + it)
        if it == 3:                       # a parameter without a gradient is skipped, like torch
            ours[4].grad = None
            ref[4].grad = None
        o.step()
        r.step()
        for i, (a, b) in enumerate(zip(ours, ref)):
            assert _rel(a.detach(), b.detach()) < 2e-6, (it, i, SHAPES[i], _rel(a.detach(), b.detach()))
    for a, b in zip(ours, ref):
        sa, sb = o.state[a], r.state[b]
        assert _rel(sa["exp_avg"], sb["exp_avg"]) < 1e-6 and _rel(sa["exp_avg_sq"], sb["exp_avg_sq"]) < 1e-6
    assert float(o.state[ours[0]]["step"]) == 6.0 and float(r.state[ref[0]]["step"]) == 6.0
    assert float(o.state[ours[4]]["step"]) == 5.0          # skipped once: the counter is per parameter, like torch's
    v0 = ours[0]._version
    _set_grads(ours, 999)
    o.step()
    assert ours[0]._version > v0                           # raw-pointer update is visible to autograd's version counters


def test_fused_adam_state_dict_round_trips_with_torch_adam(cuda):
    """The checkpoint format is torch's: a FusedAdam state loads into torch.optim.Adam and back, and the trajectories stay
    together afterwards."""
    from gif_b200.optim import FusedAdam
    a, b, c = _params(cuda, 2), _params(cuda, 2), _params(cuda, 2)
    fa = FusedAdam(a, lr=1e-3, betas=(0.0, 0.99))
    for it in range(3):
        _set_grads(a, 200 + it)
        fa.step()
    tb = torch.optim.Adam(b, lr=1e-3, betas=(0.0, 0.99), foreach=False)
    import copy
    sd = copy.deepcopy(fa.state_dict())                  # what a file round trip gives (load_state_dict itself aliases tensors)
    sd["param_groups"][0]["capturable"] = False          # the reference's optimiser is a plain one
    tb.load_state_dict(sd)
    fc = FusedAdam(c, lr=1e-3, betas=(0.0, 0.99))
    fc.load_state_dict(copy.deepcopy(tb.state_dict()))
    with torch.no_grad():
        for x, y, z in zip(a, b, c):
            y.copy_(x)
            z.copy_(x)
    for it in range(3):
        for ps in (a, b, c):
            _set_grads(ps, 300 + it)
        fa.step()
        tb.step()
        fc.step()
    for x, y, z in zip(a, b, c):
        assert _rel(y.detach(), x.detach()) < 2e-6 and _rel(z.detach(), x.detach()) < 2e-6
    assert float(fc.state[c[0]]["step"]) == 6.0


def test_fused_adam_in_a_cuda_graph(cuda):
    from gif_b200.optim import FusedAdam
    ours, ref = _params(cuda, 3), _params(cuda, 3)
    o = FusedAdam(ours, lr=1e-3, betas=(0.0, 0.99))
    r = torch.optim.Adam(ref, lr=1e-3, betas=(0.0, 0.99), foreach=False)
    _set_grads(ours, 400)
    _set_grads(ref, 400)
    static = [p.grad for p in ours]
    o.step()                                # warm-up (state allocation) outside the capture
    r.step()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            o.step()
    torch.cuda.current_stream().wait_stream(side)
    # capture does not execute: ours is one step behind until the first replay
    for it in range(3):
        _set_grads(ref, 500 + it)
        for s, p in zip(static, ref):
            s.copy_(p.grad)
        graph.replay()
        r.step()
    torch.cuda.synchronize()
    for a, b in zip(ours, ref):
        assert _rel(a.detach(), b.detach()) < 2e-6
    assert float(o.state[ours[0]]["step"]) == 4.0


def test_fused_adam_refuses_cpu_parameters():
    from gif_b200.optim import FusedAdam
    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([p]).step()
