"""The body of the reference's training loop (train.py:80-252) on the gif_b200 modules: one D step + one G step per
iteration, R1 every 16th iteration (train.py:145-149), optional path-length regularisation (train.py:205-208, adopted
rule), Adam with the reference's hyper-parameters (train.py:365-382), EMA of the generator (train.py:250).

Differences from the reference, all deliberate:
  * nn.DataParallel -> FlatGradAllReducer (one all-reduce per net per step);
  * the fake images of the D step are generated under no_grad (the reference builds, then discards, a G graph because
    gen_in1 requires grad, train.py:140,152,160 -- the detached result is identical);
  * data loading / FID / checkpoints are outside the hot path.
"""
import copy
import math

import torch
import torch.nn.functional as F

from . import losses
from .distributed import FlatGradAllReducer
from .model.stg2_discriminator import Discriminator
from .model.stg2_generator import StyledGenerator


def requires_grad(model, flag=True):
    """my_utils/generic_utils.py:58-60."""
    for p in model.parameters():
        p.requires_grad = flag


def accumulate(model1, model2, decay=0.999):
    """my_utils/generic_utils.py:63-76 (EMA), as one fused foreach update."""
    p1 = [p.data for p in model1.parameters()]
    p2 = [p.data for p in model2.parameters()]
    torch._foreach_mul_(p1, decay)
    torch._foreach_add_(p1, p2, alpha=1 - decay)


class GifTrainer:
    def __init__(self, device, resolution=256, vocab=70_000, r1_every=16, ppl=False, world_size=1, seed=0):
        torch.manual_seed(seed)
        self.device = device
        self.step_idx = int(math.log2(resolution)) - 2                           # train.py:387
        self.generator = StyledGenerator(embedding_vocab_size=vocab, rendered_flame_ascondition=True,
                                         normal_maps_as_cond=True, core_tensor_res=4, n_mlp=8).to(device)
        self.discriminator = Discriminator(resolution, num_color_chnls=9, channel_multiplier=2).to(device)
        self.g_running = copy.deepcopy(self.generator).train(False)
        g_ratio, d_ratio = 4 / 5, 16 / 17                                        # train.py:365-366
        cap = device.type == "cuda"     # capturable Adam keeps its step counter on the device (needed for CUDA graphs)
        self.g_optimizer = torch.optim.Adam(self.generator.parameters(), lr=0.002 * g_ratio, betas=(0.0, 0.99 ** g_ratio),
                                            capturable=cap)
        self.d_optimizer = torch.optim.Adam(self.discriminator.parameters(), lr=0.002 * d_ratio,
                                            betas=(0.0, 0.99 ** d_ratio), capturable=cap)
        self.g_reducer = FlatGradAllReducer(list(self.generator.parameters()), world_size)
        self.d_reducer = FlatGradAllReducer(list(self.discriminator.parameters()), world_size)
        self.r1_every = r1_every
        self.ppl = losses.PathLengthRegularizor() if ppl else None
        self.iteration = 0
        self._graphs = None
        requires_grad(self.generator, False)

    # ------------------------------------------------------------------------------------------- CUDA graphs
    def capture(self, batch, resolution):
        """Capture the whole iteration (D step + G step, both optimiser updates, EMA, the gradient all-reduces) into two
        CUDA graphs -- one for the iterations with the R1 penalty, one for those without -- so that a step is two
        cudaGraphLaunch calls instead of ~2000 kernel launches issued from Python.  Call after a few eager warm-up
        iterations (optimiser state and workspaces must exist).  ``train_iteration`` then copies its inputs into the
        static buffers and replays."""
        dev = self.device
        self._static = (torch.zeros(batch, 3, resolution, resolution, device=dev),
                        torch.zeros(batch, 6, resolution, resolution, device=dev),
                        torch.zeros(batch, dtype=torch.long, device=dev))
        from . import _lib
        graphs = {}
        pool = None
        self.graph_launches = {}
        torch.cuda.synchronize()
        for with_r1 in (True, False):
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g, pool=pool):
                out = self._iteration_body(*self._static, with_r1=with_r1)
            self.graph_launches[with_r1] = _lib.launch_count() - n0        # gif_b200 kernel nodes per replay
            pool = g.pool()
            graphs[with_r1] = (g, out)
        self._graphs = graphs

    def train_iteration(self, real_image, flm_rndr, input_indices):
        """real_image (B,3,R,R) in [-1,1], flm_rndr (B,6,R,R) in [-1,1], input_indices (B,) int64 -- device tensors.
        Returns (d_loss, g_loss) as 0-d device tensors."""
        with_r1 = (self.iteration + 1) % self.r1_every == 0                       # train.py:145
        self.iteration += 1
        if self._graphs is not None:
            self.replayed_launches = getattr(self, "replayed_launches", 0) + self.graph_launches[with_r1]
            for dst, src in zip(self._static, (real_image, flm_rndr, input_indices)):
                dst.copy_(src, non_blocking=True)
            g, out = self._graphs[with_r1]
            g.replay()
            return out
        return self._iteration_body(real_image, flm_rndr, input_indices, with_r1)

    def _iteration_body(self, real_image, flm_rndr, input_indices, with_r1):
        G, D, step = self.generator, self.discriminator, self.step_idx
        # ------------------------------------------------ D step (train.py:82-178)
        requires_grad(D, True)
        self.d_reducer.zero()
        real_image = real_image.detach().requires_grad_(True)                     # train.py:135-136
        real_scores, _ = D([real_image], condition=flm_rndr, step=step, alpha=1)
        real_loss = F.softplus(-real_scores).mean()
        if with_r1:                                                               # train.py:145-149
            real_loss = real_loss + losses.grad_penalty_loss([real_image], real_scores, step=None).mean()
        with torch.no_grad():
            fake = G(flm_rndr, None, step=step, alpha=1, input_indices=input_indices)[0]
        fake_scores, _ = D([fake], condition=flm_rndr, step=step, alpha=1)
        d_loss = real_loss + F.softplus(fake_scores).mean()
        d_loss.backward()
        self.d_reducer.all_reduce_mean()
        self.d_optimizer.step()
        # ------------------------------------------------ G step (train.py:181-252)
        requires_grad(G, True)
        requires_grad(D, False)
        self.g_reducer.zero()
        fake = G(flm_rndr, None, step=step, alpha=1, input_indices=input_indices)[0]
        predict, _ = D([fake], condition=flm_rndr, step=step, alpha=1)
        g_loss = F.softplus(-predict).mean()
        if self.ppl is not None:
            g_loss = g_loss + 2 * self.ppl.path_length_reg(G, step=step, alpha=1, input_indices=input_indices,
                                                           cond=flm_rndr)       # train.py:205-208
        g_loss.backward()
        self.g_reducer.all_reduce_mean()
        self.g_optimizer.step()
        accumulate(self.g_running, G, decay=0.5 ** (32 / (10 * 1000)))            # train.py:250
        requires_grad(G, False)
        return d_loss.detach(), g_loss.detach()
