"""StyleGAN2 residual discriminator with the API of the reference's model/stg2_discriminator.py ("disc.py"): same
constructor / forward signature, same state_dict keys (convs.{i}..., final_conv..., final_linear.{0,1}...), computed
channels-last through the gif_b200 kernels."""
import math

import torch
from torch import nn

from .. import ops
from .stylegan2_common_layers import ConvLayer, EqualLinear, ResBlock


def _channel_table(multiplier):
    """Feature widths per resolution (disc.py:12-22): 512 up to 32^2, then 256m, 128m, ... halving per octave."""
    table = {res: 512 for res in (4, 8, 16, 32)}
    width = 256
    for res in (64, 128, 256, 512, 1024):
        table[res] = width * multiplier
        width //= 2
    return table


def minibatch_stddev_feature(feat, group_size=4, n_feat=1):
    """disc.py:59-65 on an NCHW tensor: standard deviation over groups of ``group_size`` samples (biased variance, +1e-8),
    averaged over channels and pixels, appended as ``n_feat`` extra constant channels.  Tiny ((B,512,4,4)): torch glue."""
    b, c, h, w = feat.shape
    g = min(b, group_size)
    grouped = feat.reshape(g, b // g, n_feat, c // n_feat, h, w)
    sd = (grouped.var(dim=0, unbiased=False) + 1e-8).sqrt()              # (b/g, n_feat, c/n_feat, h, w)
    sd = sd.mean(dim=(2, 3, 4))                                          # (b/g, n_feat)
    plane = sd[:, :, None, None].repeat(g, 1, h, w)                      # sample i of every group gets its group's value
    return torch.cat([feat, plane], dim=1)


class Discriminator(nn.Module):
    """disc.py:8-76."""

    def __init__(self, size, channel_multiplier=2, num_color_chnls=3, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        widths = _channel_table(channel_multiplier)
        octaves = int(math.log2(size))
        resolutions = [2 ** o for o in range(octaves, 1, -1)]           # size, size/2, ..., 4
        stem = ConvLayer(num_color_chnls, widths[size], 1)
        blocks = [ResBlock(widths[hi], widths[lo], blur_kernel) for hi, lo in zip(resolutions[:-1], resolutions[1:])]
        self.convs = nn.Sequential(stem, *blocks)
        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(widths[4] + 1, widths[4], 3)
        self.final_linear = nn.Sequential(EqualLinear(widths[4] * 16, widths[4], activation='fused_lrelu'),
                                          EqualLinear(widths[4], 1))

    def forward(self, input, condition=None, step=0, alpha=0):
        if type(input) in (list, tuple):
            input = input[0]
        if condition is not None:
            input = torch.cat((input, condition), axis=1)                 # disc.py:53 (9 channels)
        x = ops.from_reference_nchw(input)           # gradients w.r.t. the caller's images come back NCHW-contiguous
        x = self.convs[0].forward_nhwc(x, rt_out=True)
        for block in list(self.convs)[1:]:
            x = block.forward_nhwc(x)
        out = minibatch_stddev_feature(ops.to_nchw_view(x), self.stddev_group, self.stddev_feat)
        batch = out.shape[0]
        x = self.final_conv.forward_nhwc(ops.to_nhwc(out))
        out = ops.to_nchw_view(x).reshape(batch, -1)                      # NCHW flatten order (disc.py:70)
        out = self.final_linear(out)
        return out, None
