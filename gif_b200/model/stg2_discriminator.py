"""StyleGAN2 residual discriminator with the API of the reference's model/stg2_discriminator.py ("disc.py")."""
import math

import torch
from torch import nn

from .. import ops
from .stylegan2_common_layers import ConvLayer, EqualLinear, ResBlock


class Discriminator(nn.Module):
    """disc.py:8-76."""

    def __init__(self, size, channel_multiplier=2, num_color_chnls=3, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        m = channel_multiplier
        channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * m, 128: 128 * m, 256: 64 * m, 512: 32 * m,
                    1024: 16 * m}
        convs = [ConvLayer(num_color_chnls, channels[size], 1)]
        log_size = int(math.log(size, 2))
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(
            EqualLinear(channels[4] * 4 * 4, channels[4], activation='fused_lrelu'),
            EqualLinear(channels[4], 1),
        )

    def forward(self, input, condition=None, step=0, alpha=0):
        if type(input) in (list, tuple):
            input = input[0]
        if condition is not None:
            input = torch.cat((input, condition), axis=1)                 # disc.py:53 (9 channels)
        x = ops.to_nhwc(input)
        x = self.convs[0].forward_nhwc(x, rt_out=True)
        for block in list(self.convs)[1:]:
            x = block.forward_nhwc(x)
        # minibatch standard deviation (disc.py:59-65) on the (B,4,4,512) tail: tiny, torch glue.
        out = ops.to_nchw_view(x)
        batch, channel, height, width = out.shape
        group = min(batch, self.stddev_group)
        stddev = out.reshape(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        stddev = torch.sqrt(stddev.var(0, unbiased=False) + 1e-8)
        stddev = stddev.mean([2, 3, 4], keepdims=True).squeeze(2)
        stddev = stddev.repeat(group, 1, height, width)
        out = torch.cat([out, stddev], 1)
        x = self.final_conv.forward_nhwc(ops.to_nhwc(out))
        out = ops.to_nchw_view(x).reshape(batch, -1)                      # NCHW flatten order (disc.py:70)
        out = self.final_linear(out)
        return out, None
