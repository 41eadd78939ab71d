"""CPU, world_size 2, gloo: the gradient exchange that replaces nn.DataParallel (gif_b200/distributed.py).
Host-side logic only (flat buffer, .grad views, requires_grad toggling, unused parameters, replicas stay identical)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from gif_b200.distributed import FlatGradAllReducer, broadcast_module, init_from_env
    r, w, _ = init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                       # replicas start different ...
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    unused = torch.nn.Parameter(torch.randn(5))          # like progression.7/8, to_rgb.7/8 at step 6
    broadcast_module(net)                                # ... and are made identical
    dist.broadcast(unused.data, 0)
    params = list(net.parameters()) + [unused]
    red = FlatGradAllReducer(params, world)
    opt = torch.optim.Adam(params, lr=1e-2, betas=(0.0, 0.99))
    ref = [p.detach().clone() for p in params]
    for it in range(3):
        red.zero()
        x = torch.randn(6, 8, generator=torch.Generator().manual_seed(1000 * it + rank))   # per-rank batch
        loss = net(x).pow(2).mean()
        loss.backward()
        local = [p.grad.detach().clone() for p in net.parameters()]
        red.all_reduce_mean()
        # averaged gradient == mean over ranks of the local gradients
        gathered = [torch.zeros_like(torch.cat([g.reshape(-1) for g in local])) for _ in range(world)]
        dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
        mean = sum(gathered) / world
        got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        assert torch.allclose(got, mean, atol=1e-7), (got - mean).abs().max()
        assert torch.count_nonzero(unused.grad) == 0
        opt.step()
    # replicas identical after 3 steps
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    assert torch.equal(both[0], both[1])
    assert not torch.equal(flat, torch.cat([p.reshape(-1) for p in ref]))
    # a .grad replaced by autograd (not a view of the flat buffer) is folded back in
    net[0].weight.grad = torch.ones_like(net[0].weight) * (rank + 1)
    red.all_reduce_mean()
    assert torch.allclose(net[0].weight.grad, torch.full_like(net[0].weight, 1.5))
    dist.destroy_process_group()
    q.put(rank)


def test_flat_grad_allreduce_gloo_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]


def _shim_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from gif_b200 import distributed as D
    D.init_from_env("gloo")
    original = D.install_data_parallel_shim()
    try:
        assert torch.nn.DataParallel is D.DataParallel and original is not D.DataParallel
        torch.manual_seed(7)                                   # identical replicas, as after loading one checkpoint
        net = torch.nn.DataParallel(torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4)))
        unused = torch.nn.Parameter(torch.randn(3))             # never receives a gradient (train.py: progression.7/8)
        assert list(net.state_dict())[0] == "module.0.weight"   # checkpoint keys keep the reference's prefix
        opt = torch.optim.Adam(list(net.module.parameters()) + [unused], lr=1e-2, betas=(0.0, 0.99))   # train.py:367
        for it in range(3):
            net.zero_grad()
            x = torch.randn(6, 8, generator=torch.Generator().manual_seed(50 * it + rank))   # per-rank batch
            net(x).pow(2).mean().backward()
            local = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
            opt.step()                                          # the hook averages the gradients before the update
            both = [torch.zeros_like(local) for _ in range(world)]
            dist.all_gather(both, local)
            got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
            assert torch.allclose(got, sum(both) / world, atol=1e-7)
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        every = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(every, flat)
        assert torch.equal(every[0], every[1])                  # replicas stay bit-identical
    finally:
        D.uninstall_data_parallel_shim()
    assert torch.nn.DataParallel is original
    dist.destroy_process_group()
    q.put(rank)


def test_data_parallel_shim_and_optimizer_hook_gloo_world2():
    """train.py:344-367 unchanged under one process per GPU: nn.DataParallel stand-in + gradient all-reduce in the
    optimiser pre-step hook."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shim_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]
