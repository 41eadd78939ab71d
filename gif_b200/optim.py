"""Adam for the loop body (train.py:365-382) as multi-tensor CUDA launches.

``FusedAdam`` IS a ``torch.optim.Adam`` (same constructor arguments the reference passes, same ``param_groups`` /
``state`` layout -- ``step``, ``exp_avg``, ``exp_avg_sq`` per parameter -- so ``state_dict()`` / ``load_state_dict()`` and
the reference's checkpoint format are unchanged); only ``step()`` is replaced: one ``gifb200_adam_step`` call per parameter
group (two launches per 64 tensors) instead of torch's capturable foreach path, whose per-parameter 0-d step counters cost
~600 tiny kernels per optimiser step.  The step counters stay per parameter (0-d device floats, incremented on the device:
CUDA-graph replays advance them; a parameter without a gradient keeps its count, like torch).  No CPU fallback: CUDA fp32
parameters only.
"""
import ctypes

import torch

from ._lib import check, lib, stream


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, capturable=True, foreach=False)
        self._tables = {}

    def _init_state(self, group):
        """Lazy state init in torch's layout (``step``: a 0-d fp32 device tensor per parameter, like capturable=True)."""
        for p in group["params"]:
            if p.grad is None:
                continue
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError("FusedAdam: contiguous CUDA float32 parameters only (no CPU fallback)")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            elif not (torch.is_tensor(st["step"]) and st["step"].is_cuda and st["step"].dtype == torch.float32):
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("FusedAdam: optimiser state was replaced during CUDA-graph capture")
                st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32).to(p.device)   # a loaded host counter

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise RuntimeError("FusedAdam: weight_decay / amsgrad / maximize are not part of the reference's configuration")
            self._init_state(group)
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                g = p.grad
                if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and not g.is_sparse):
                    raise RuntimeError("FusedAdam: contiguous dense CUDA float32 gradients only")
            n = len(ps)
            key = (tuple(p.data_ptr() for p in ps), tuple(p.grad.data_ptr() for p in ps),
                   tuple(self.state[p]["exp_avg"].data_ptr() for p in ps), tuple(self.state[p]["exp_avg_sq"].data_ptr() for p in ps),
                   tuple(self.state[p]["step"].data_ptr() for p in ps))
            tab = self._tables.get(gi)
            if tab is None or tab[0] != key:
                arr = ctypes.c_void_p * n
                tab = (key, arr(*[p.data_ptr() for p in ps]), arr(*[p.grad.data_ptr() for p in ps]),
                       arr(*[self.state[p]["exp_avg"].data_ptr() for p in ps]),
                       arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in ps]),
                       arr(*[self.state[p]["step"].data_ptr() for p in ps]),
                       (ctypes.c_longlong * n)(*[p.numel() for p in ps]))
                self._tables[gi] = tab
            beta1, beta2 = group["betas"]
            lr = group["lr"]
            if torch.is_tensor(lr):
                lr = float(lr)
            if len(set(key[4])) != n:
                raise RuntimeError("FusedAdam: two parameters share one step counter (aliased optimiser state)")
            check(lib.gifb200_adam_step(tab[1], tab[2], tab[3], tab[4], tab[5], tab[6], n, float(lr), float(beta1), float(beta2),
                                        float(group["eps"]), stream()), "gifb200_adam_step")
            # the kernel wrote the parameters through raw pointers: tell autograd (and everything keyed on a parameter's
            # version, e.g. ops.prep_weight's cache of prepared weights) that they changed, like torch's in-place ops do
            torch.autograd.graph.increment_version(ps)
        return loss
