"""The StyleGAN3-style discriminator of the reference on the libsg2b200 ops.

Drop-in for the discriminator half of implementations/StyleGAN3/model.py (:382-510): ``binomial_filter``, ``Linear``
(:16-30), ``ConvAct`` (:389-417), ``ResBlock`` (:419-440), ``MinibatchStdDev`` (:442-462), ``DiscEpilogue`` (:464-479),
``Discriminator`` (:481-510) -- same constructor arguments, attribute tree and ``state_dict`` keys.  It is the network the
reference also uses for ADA / APA / CIPS; here it exercises the generic zero-pad ``upfirdn2d`` (4x4 binomial blur in front
of every down-sampling convolution), ``conv2d_resample``, and ``bias_act`` with the sqrt(2) gain.

State: first correct path -- parity-checked against reference-generated goldens (tests/golden/sg3d.npz).  The stride-2
convolutions run as stride-1 launches whose result is decimated (ops/conv2d_gradfix.py); the generator half of that file
(filtered_lrelu) is not built.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import bias_act, conv2d_resample


def binomial_filter(filter_size):
    """Binomial taps of the given size (reference :382-387)."""
    def c(n, k):
        if k <= 0 or n <= k:
            return 1
        return c(n - 1, k - 1) + c(n - 1, k)
    return [c(filter_size - 1, j) for j in range(filter_size)]


class Linear(nn.Module):
    def __init__(self, in_features, out_features, bias, act_name='linear', gain=1.) -> None:
        super().__init__()
        self.act_name = act_name
        self.weight = nn.Parameter(torch.randn(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        self.scale = gain / (self.weight[0].numel() ** 0.5)

    def forward(self, x):
        x = F.linear(x, self.weight * self.scale)
        return bias_act.bias_act(x, self.bias.to(x.dtype), act=self.act_name)


class ConvAct(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, down=1, filter_size=4, act_name='linear',
                 gain=1., act_gain=None) -> None:
        super().__init__()
        self.down = down
        self.act_name = act_name
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.scale = gain / (self.weight[0].numel() ** 0.5)
        self.act_gain = bias_act.activation_funcs[act_name].def_gain if act_gain is None else act_gain
        if down > 1:
            taps = torch.tensor(binomial_filter(filter_size), dtype=torch.float32)
            kernel = torch.outer(taps, taps)
            kernel /= kernel.sum()
            self.register_buffer('down_filter', kernel)
        else:
            self.down_filter = None

    def forward(self, x):
        weight = self.weight * self.scale
        x = conv2d_resample.conv2d_resample(x, weight.to(x.dtype), self.down_filter, 1, self.down, self.padding)
        b = self.bias.to(x.dtype) if self.bias is not None else self.bias
        return bias_act.bias_act(x, b, act=self.act_name, gain=self.act_gain)


class ResBlock(nn.Module):
    def __init__(self, in_channels, out_channels, filter_size=4, act_name='lrelu', gain=1.) -> None:
        super().__init__()
        self.conv1 = ConvAct(in_channels, out_channels, 3, True, 1, filter_size, act_name, gain)
        self.conv2 = ConvAct(out_channels, out_channels, 3, True, 2, filter_size, act_name, gain, 0.5 ** 0.5)
        self.skip = ConvAct(in_channels, out_channels, 1, False, 2, filter_size, 'linear', gain, 0.5 ** 0.5)

    def forward(self, x):
        h = self.conv1(x)
        h = self.conv2(h)
        x = self.skip(x)
        return h + x


class MinibatchStdDev(nn.Module):
    """Reference :442-462 (num_channels statistics per group, eps 1e-8 inside the sqrt); a [B, C, 4, 4] tensor, kept in torch."""

    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size = group_size
        self.num_channels = num_channels

    def forward(self, x):
        N, C, H, W = x.shape
        G = self.group_size if N % self.group_size == 0 else N
        Fc = self.num_channels
        c = C // Fc
        y = x.reshape(G, -1, Fc, c, H, W)
        y = y - y.mean(dim=0)
        y = y.square().mean(dim=0)
        y = (y + 1e-8).sqrt()
        y = y.mean(dim=[2, 3, 4])
        y = y.reshape(-1, Fc, 1, 1)
        y = y.repeat(G, 1, H, W)
        return torch.cat([x, y], dim=1)


class DiscEpilogue(nn.Module):
    def __init__(self, mbsd_group_size, mbsd_channels, channels, bottom, act_name='lrelu', gain=1.) -> None:
        super().__init__()
        self.epilogue = nn.Sequential(
            MinibatchStdDev(mbsd_group_size, mbsd_channels),
            ConvAct(channels + mbsd_channels, channels, 3, True, 1, None, act_name, gain),
            nn.Flatten(),
            Linear(channels * bottom ** 2, channels, True, act_name, gain),
            Linear(channels, 1, True, 'linear', gain))

    def forward(self, x):
        return self.epilogue(x)


class Discriminator(nn.Module):
    def __init__(self, image_size, in_channels=3, channels=64, max_channels=512, kernel_size=3, mbsd_group_size=4,
                 mbsd_channels=1, bottom=4, filter_size=4, act_name='lrelu', gain=1.) -> None:
        super().__init__()
        num_downs = int(math.log2(image_size) - math.log2(bottom))
        ochannels = channels
        self.from_rgb = ConvAct(in_channels, ochannels, 1, True, 1, None, act_name, gain)
        resblocks = []
        for _ in range(num_downs):
            channels *= 2
            ichannels, ochannels = ochannels, min(max_channels, channels)
            resblocks.append(ResBlock(ichannels, ochannels, filter_size, act_name, gain))
        self.resblocks = nn.Sequential(*resblocks)
        self.epilogue = DiscEpilogue(mbsd_group_size, mbsd_channels, ochannels, bottom, act_name, gain)

    def forward(self, x):
        x = self.from_rgb(x)
        x = self.resblocks(x)
        return self.epilogue(x)
