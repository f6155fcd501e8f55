"""StyleGAN3-style networks of the reference on the libsg2b200 ops.

Discriminator half of implementations/StyleGAN3/model.py (:382-510) -- the StyleGAN2-ADA discriminator the reference also
uses for ADA / APA / CIPS.  Constructor arguments, attribute names and ``state_dict`` keys are the reference's (checkpoints
load either way; that is how parity is tested, tests/golden/sg3d.npz); the execution is this package's:
  * a stride-1 ``ConvAct`` is ONE launch -- convolution + bias + leaky-ReLU + gain in the tcgen05 kernel's epilogue
    (``ops.conv2d.conv2d_bias_act``) -- instead of ``conv2d_resample`` followed by ``bias_act`` (:411-416);
  * a down-sampling ``ConvAct`` runs the 4x4 binomial low-pass on the ``upfirdn2d`` register-ring kernel, the stride-2
    convolution through ``conv2d_resample`` and the bias / gain in one ``bias_act`` pass;
  * ``Linear`` is one fused launch (``ops.linear``: scale, bias, activation and its gain), no cuBLAS;
  * ``MinibatchStdDev`` with one statistic channel is the two-launch ``ops.mbstd`` kernel (eps 1e-8 inside the square root,
    :457), any other channel count the reference's tensor expression.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .ops import bias_act as _ba
from .ops import conv2d as _conv
from .ops import conv2d_resample as _resample
from .ops.linear import linear_bias_act
from .ops.mbstd import minibatch_stddev

_LRELU_SLOPE = _ba.activation_funcs['lrelu'].def_alpha


def binomial_filter(filter_size):
    """Row `filter_size - 1` of Pascal's triangle (reference :382-387 computes the same numbers recursively)."""
    return [math.comb(filter_size - 1, j) for j in range(filter_size)]


def _fan_in_scale(weight, gain):
    return gain / math.sqrt(weight[0].numel())


def _fused_act(act_name):
    """(slope or None) when the activation can ride in a kernel epilogue, else raises KeyError."""
    return {'linear': None, 'lrelu': _LRELU_SLOPE}[act_name]


class Linear(nn.Module):
    """act(x (W * scale)^T + b) * def_gain(act)  (reference :16-30)."""

    def __init__(self, in_features, out_features, bias, act_name='linear', gain=1.) -> None:
        super().__init__()
        self.act_name = act_name
        self.weight = nn.Parameter(torch.randn(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        self.scale = _fan_in_scale(self.weight, gain)

    def forward(self, x):
        if self.act_name in ('linear', 'lrelu') and x.dtype == torch.float32:
            return linear_bias_act(x, self.weight, self.bias, self.scale, _ba.activation_funcs[self.act_name].def_gain,
                                   _fused_act(self.act_name))
        y = linear_bias_act(x.float(), self.weight, None, self.scale)
        return _ba.bias_act(y.to(x.dtype), None if self.bias is None else self.bias.to(x.dtype), act=self.act_name)


class ConvAct(nn.Module):
    """[low-pass -> stride-`down`] convolution -> bias -> activation * act_gain  (reference :389-417)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, down=1, filter_size=4, act_name='linear',
                 gain=1., act_gain=None) -> None:
        super().__init__()
        self.down = down
        self.act_name = act_name
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.scale = _fan_in_scale(self.weight, gain)
        self.act_gain = _ba.activation_funcs[act_name].def_gain if act_gain is None else act_gain
        if down > 1:
            taps = torch.tensor(binomial_filter(filter_size), dtype=torch.float32)
            self.register_buffer('down_filter', torch.outer(taps, taps) / taps.sum() ** 2)
        else:
            self.down_filter = None

    def forward(self, x):
        one_launch = (self.down == 1 and self.act_name in ('linear', 'lrelu') and x.dtype == torch.float32
                      and self.weight.shape[2] in (1, 3) and self.act_gain > 0)
        if one_launch:
            return _conv.conv2d_bias_act(x, self.weight, self.bias, self.scale, _fused_act(self.act_name), self.act_gain)
        y = _resample.conv2d_resample(x, (self.weight * self.scale).to(x.dtype), self.down_filter, 1, self.down, self.padding)
        b = None if self.bias is None else self.bias.to(y.dtype)
        return _ba.bias_act(y, b, act=self.act_name, gain=self.act_gain)


class ResBlock(nn.Module):
    """conv1 -> down-sampling conv2, plus a down-sampling 1x1 skip; both branches scaled by sqrt(1/2) (reference :419-440)."""

    def __init__(self, in_channels, out_channels, filter_size=4, act_name='lrelu', gain=1.) -> None:
        super().__init__()
        half = math.sqrt(0.5)
        self.conv1 = ConvAct(in_channels, out_channels, 3, True, 1, filter_size, act_name, gain)
        self.conv2 = ConvAct(out_channels, out_channels, 3, True, 2, filter_size, act_name, gain, half)
        self.skip = ConvAct(in_channels, out_channels, 1, False, 2, filter_size, 'linear', gain, half)

    def forward(self, x):
        return self.conv2(self.conv1(x)) + self.skip(x)


class MinibatchStdDev(nn.Module):
    """Appends `num_channels` group-statistics channels (reference :442-462)."""

    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size = group_size
        self.num_channels = num_channels

    def forward(self, x):
        if self.num_channels == 1 and x.is_cuda and x.dtype == torch.float32:
            return minibatch_stddev(x, self.group_size, 1e-8)
        n, c, h, w = x.shape
        groups = self.group_size if n % self.group_size == 0 else n
        f = self.num_channels
        dev = x.reshape(groups, -1, f, c // f, h, w)
        dev = dev - dev.mean(0)
        stat = dev.square().mean(0).add(1e-8).sqrt().mean([2, 3, 4])            # [n / groups, f]
        return torch.cat([x, stat.reshape(-1, f, 1, 1).repeat(groups, 1, h, w)], 1)


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
        widths = [channels]
        for _ in range(int(math.log2(image_size) - math.log2(bottom))):
            channels *= 2
            widths.append(min(max_channels, channels))
        self.from_rgb = ConvAct(in_channels, widths[0], 1, True, 1, None, act_name, gain)
        self.resblocks = nn.Sequential(*[ResBlock(a, b, filter_size, act_name, gain) for a, b in zip(widths[:-1], widths[1:])])
        self.epilogue = DiscEpilogue(mbsd_group_size, mbsd_channels, widths[-1], bottom, act_name, gain)

    def forward(self, x):
        return self.epilogue(self.resblocks(self.from_rgb(x)))
