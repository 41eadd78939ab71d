"""One-process-per-GPU data parallelism: replaces train.py's nn.DataParallel (train.py:344,356,358).

The reference re-broadcasts every parameter on each of its 5 wrapped forwards per iteration and reduces gradients to
GPU 0.  Here every rank holds full replicas (G, D, EMA-G, Adam state), draws its own batch, and after each
``backward()`` the gradients of the net being optimised are averaged with ONE collective per net on a pre-flattened fp32
buffer (``torch.distributed`` all-reduce: NCCL over NVLink 5 / NVSwitch on GPUs, gloo in the CPU tests).  Replicas stay
bit-identical because they apply identical averaged gradients.  Minibatch-stddev groups stay inside a rank
(per-rank batch % 4 == 0), the rasteriser is per-sample: no other exchange exists on the path.

A plain DistributedDataParallel wrapper does not fit the reference's loop (requires_grad toggling per phase,
``autograd.grad`` for R1, two forwards per backward, parameters of unused resolutions), hence the explicit buffer.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


class FlatGradAllReducer:
    """Gradient exchange for one network: a flat fp32 buffer with one slot per parameter that can receive a gradient."""

    def __init__(self, params, world_size=None, group=None):
        self.params = [p for p in params]
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        self.nbytes = n * 4

    def attach(self):
        """Make every parameter's .grad a view into the flat buffer (so backward accumulates straight into it)."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()
        for p, v in zip(self.params, self.views):   # parameters unused in this step keep a zero gradient
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def all_reduce_mean(self):
        """SUM over ranks then 1/world.  Parameters whose .grad was replaced by autograd are copied back first."""
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)


def broadcast_module(module, src=0):
    """Make all replicas start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)
