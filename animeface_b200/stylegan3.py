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


# =====================================================================================================================
# Generator half (implementations/StyleGAN3/model.py:32-380; SURVEY 8f n3, BASELINE config 5).  Attribute names, buffers and
# ``state_dict`` keys are the reference's.  Execution:
#   * ``ModulatedConv`` is the library's fused modulated convolution (style and the magnitude-EMA input gain scale the
#     activation tile, demodulation is an epilogue scale): the reference's per-sample weight tensor [B, Co, Ci, k, k] and its
#     ``groups = B`` convolution (:53-72) are never built;
#   * the alias-free non-linearity is ``ops.filtered_lrelu`` (bias, x`up` FIR, leaky ReLU * gain + clamp, FIR /`down`);
#   * ``SynthesisInput``'s trainable channel mixing (:266) is a 1x1 convolution on the library kernels;
#   * ``Linear`` / ``PixelNorm`` are the fused kernels of ``ops.linear``.
import numpy as np

from .ops.filtered_lrelu import filtered_lrelu
from .ops.linear import pixel_norm


class ModulatedConv(nn.Module):
    """Weight-(de)modulated convolution with 'full' padding k - 1 (reference :32-72)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1, demod=True) -> None:
        super().__init__()
        self.in_channels = in_channels
        self.padding = padding
        self.demod = demod
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(self.weight[0].numel())

    def forward(self, x, s, input_gain=None):
        k = self.weight.shape[2]
        extra = self.padding - k // 2                # the kernels convolve 'same': a wider padding is applied to the input
        if extra > 0:
            x = torch.nn.functional.pad(x, [extra] * 4)
        gain = None if input_gain is None else input_gain.expand(x.shape[0], self.in_channels)
        y = _conv.modulated_conv2d(x, self.weight, s, None, None, self.demod, None, eps=1e-8, in_gain=gain)
        return y if extra >= 0 else y[:, :, -extra:y.shape[2] + extra, -extra:y.shape[3] + extra]


def design_filter(numtaps, cutoff, width, fs, radial=False):
    """Kaiser-windowed low-pass: separable sinc (scipy.signal.firwin) or, radial, a jinc filter with a separable Kaiser
    window, normalised to unit DC gain (reference :74-90)."""
    import scipy.signal
    import scipy.special
    assert numtaps >= 1
    if numtaps == 1:
        return None
    if not radial:
        return torch.as_tensor(scipy.signal.firwin(numtaps=numtaps, cutoff=cutoff, width=width, fs=fs), dtype=torch.float32)
    pos = (np.arange(numtaps) - (numtaps - 1) / 2) / fs
    radius = np.hypot(*np.meshgrid(pos, pos))
    taps = scipy.special.j1(2 * cutoff * (np.pi * radius)) / (np.pi * radius)
    window = np.kaiser(numtaps, scipy.signal.kaiser_beta(scipy.signal.kaiser_atten(numtaps, width / (fs / 2))))
    taps = taps * np.outer(window, window)
    return torch.as_tensor(taps / np.sum(taps), dtype=torch.float32)


def get_layer_params(image_size, num_layers, channels, max_channels=512, image_channels=3, margin_size=10,
                     first_cutoff=2, first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, num_critical=2):
    """Per-layer channels, sizes, sampling rates, cutoffs and transition half-widths: cutoff and stop band grow geometrically
    from the first layer to the image's Nyquist-limited last ones (reference :92-113)."""
    last_cutoff = image_size / 2
    last_stopband = last_cutoff * last_stopband_rel
    t = np.minimum(np.arange(num_layers + 1) / (num_layers - num_critical), 1)
    cutoffs = first_cutoff * (last_cutoff / first_cutoff) ** t
    stopbands = first_stopband * (last_stopband / first_stopband) ** t
    sampling_rates = np.exp2(np.ceil(np.log2(np.minimum(stopbands * 2, image_size))))
    half_widths = np.maximum(stopbands, sampling_rates / 2) - cutoffs
    sizes = sampling_rates + margin_size * 2
    sizes[-2:] = image_size
    widths = np.rint(np.minimum((channels / 2) / cutoffs, max_channels))
    widths[-1] = image_channels
    return widths, sizes, sampling_rates, cutoffs, half_widths


class StyleLayer(nn.Module):
    """affine -> modulated conv (input scaled by the running magnitude) -> filtered leaky ReLU (reference :115-193)."""

    def __init__(self, in_channels, style_dim, out_channels, kernel_size, in_size, out_size, in_sampling_rate, out_sampling_rate,
                 in_cutoff, out_cutoff, in_half_width, out_half_width, is_rgb, is_critical_sampled,
                 lrelu_sampling=2, filter_size=6, conv_clamp=256, ema_decay=0.999) -> None:
        super().__init__()
        self.conv_clamp = conv_clamp
        self.ema_decay = ema_decay
        self.is_rgb = is_rgb
        self.gain = 1. if is_rgb else math.sqrt(2)
        self.negative_slope = 1. if is_rgb else 0.2
        self.affine = Linear(style_dim, in_channels, True)
        self.affine.bias.data.fill_(1.)
        self.register_buffer('ema', torch.ones([]))
        # the non-linearity runs at `lrelu_sampling` x the faster of the two rates; filters for getting there and back
        work_rate = max(in_sampling_rate, out_sampling_rate) * (1 if is_rgb else lrelu_sampling)
        self.up_factor = int(np.rint(work_rate / in_sampling_rate))
        self.down_factor = int(np.rint(work_rate / out_sampling_rate))
        assert in_sampling_rate * self.up_factor == work_rate and out_sampling_rate * self.down_factor == work_rate
        up_taps = filter_size * self.up_factor if self.up_factor > 1 and not is_rgb else 1
        down_taps = filter_size * self.down_factor if self.down_factor > 1 and not is_rgb else 1
        self.register_buffer('up_filter', design_filter(up_taps, in_cutoff, in_half_width * 2, work_rate))
        self.register_buffer('down_filter', design_filter(down_taps, out_cutoff, out_half_width * 2, work_rate, not is_critical_sampled))
        # padding (in up-sampled pixels) that makes the layer map in_size -> out_size exactly
        in_size = np.broadcast_to(np.asarray(in_size), [2])
        out_size = np.broadcast_to(np.asarray(out_size), [2])
        total = (out_size - 1) * self.down_factor + 1 - (in_size + kernel_size - 1) * self.up_factor + up_taps + down_taps - 2
        lo = (total + self.up_factor) // 2
        hi = total - lo
        self.padding = [int(lo[0]), int(hi[0]), int(lo[1]), int(hi[1])]
        self.conv = ModulatedConv(in_channels, out_channels, kernel_size, kernel_size - 1, not is_rgb)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, w):
        if self.training:
            power = x.detach().to(torch.float32).square().mean()
            self.ema.copy_(power.lerp_(self.ema, self.ema_decay))
        x = self.conv(x, self.affine(w), self.ema.rsqrt())
        return filtered_lrelu(x, self.up_filter, self.down_filter, self.bias.to(x.dtype), self.up_factor, self.down_factor,
                              self.padding, self.gain, self.negative_slope, self.conv_clamp)


class SynthesisInput(nn.Module):
    """Fourier features under a style-predicted rotation + translation, mixed by a trainable matrix (reference :195-269)."""

    def __init__(self, style_dim, channels, size, sampling_rate, bandwidth) -> None:
        super().__init__()
        self.channels = channels
        self.bandwidth = bandwidth
        self.sampling_rate = sampling_rate
        self.size = [int(v) for v in np.broadcast_to(np.asarray(size), [2])]
        freqs = torch.randn(channels, 2)
        radii = freqs.square().sum(1, keepdim=True).sqrt()
        freqs /= radii * radii.square().exp().pow(0.25)
        freqs *= bandwidth
        phases = torch.rand(channels) - 0.5
        self.weight = nn.Parameter(torch.randn(channels, channels))
        self.scale = 1 / math.sqrt(channels)
        self.affine = Linear(style_dim, 4, True)
        self.affine.weight.data.fill_(0.)                                   # identity transform by default
        self.affine.bias.data.copy_(torch.tensor([1, 0, 0, 0], dtype=torch.float32))
        self.register_buffer('transform', torch.eye(3, 3))
        self.register_buffer('freqs', freqs)
        self.register_buffer('phases', phases)

    def forward(self, w):
        B, dev = w.size(0), w.device
        t = self.affine(w)
        t = t / t[:, :2].norm(dim=1, keepdim=True)                          # (cos, sin, tx, ty)
        zero, one = torch.zeros_like(t[:, 0]), torch.ones_like(t[:, 0])
        rot = torch.stack([torch.stack([t[:, 0], -t[:, 1], zero], 1), torch.stack([t[:, 1], t[:, 0], zero], 1), torch.stack([zero, zero, one], 1)], 1)
        shift = torch.stack([torch.stack([one, zero, -t[:, 2]], 1), torch.stack([zero, one, -t[:, 3]], 1), torch.stack([zero, zero, one], 1)], 1)
        m = rot @ shift @ self.transform.unsqueeze(0)
        phases = self.phases.unsqueeze(0) + (self.freqs.unsqueeze(0) @ m[:, :2, 2:]).squeeze(2)
        freqs = self.freqs.unsqueeze(0) @ m[:, :2, :2]
        amp = (1 - (freqs.norm(dim=2) - self.bandwidth) / (self.sampling_rate / 2 - self.bandwidth)).clamp(0, 1)   # fade out-of-band
        theta = torch.eye(2, 3, device=dev)
        theta[0, 0] = 0.5 * self.size[0] / self.sampling_rate
        theta[1, 1] = 0.5 * self.size[1] / self.sampling_rate
        grid = torch.nn.functional.affine_grid(theta.unsqueeze(0), [1, 1, self.size[1], self.size[0]], align_corners=False)
        x = (grid.unsqueeze(3) @ freqs.permute(0, 2, 1).unsqueeze(1).unsqueeze(2)).squeeze(3)          # [B, H, W, C]
        x = torch.sin((x + phases.unsqueeze(1).unsqueeze(2)) * (np.pi * 2)) * amp.unsqueeze(1).unsqueeze(2)
        # trainable channel mixing = a 1x1 convolution with weight * scale
        return _conv.conv2d(x.permute(0, 3, 1, 2), self.weight[:, :, None, None], self.scale)


class PixelNorm(nn.Module):
    def forward(self, x):
        return pixel_norm(x, 1e-8)


class Mapping(nn.Module):
    def __init__(self, latent_dim, style_dim, num_layers=2, pixel_norm=True, ema_decay=0.998) -> None:
        super().__init__()
        self.ema_decay = ema_decay
        if pixel_norm:
            self.norm = PixelNorm()
        self.net = nn.Sequential(*[Linear(latent_dim if i == 0 else style_dim, style_dim, True, 'lrelu') for i in range(num_layers)])
        self.register_buffer('w_avg', torch.zeros(style_dim))

    def forward(self, z, truncation_psi=1.):
        if hasattr(self, 'norm'):
            z = self.norm(z)
        w = self.net(z)
        if self.training:
            self.w_avg.copy_(w.detach().to(torch.float32).mean(dim=0).lerp(self.w_avg, self.ema_decay))
        return w if truncation_psi == 1 else self.w_avg.lerp(w, truncation_psi)


class Synthesis(nn.Module):
    def __init__(self, image_size, num_layers=14, channels=32, max_channels=512, style_dim=512, image_channels=3, output_scale=0.25,
                 margin_size=10, first_cutoff=2, first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, kernel_size=3) -> None:
        super().__init__()
        self.num_ws = num_layers + 2                    # input + layers + ToRGB
        # the reference's width convention (:318-325): `channels` is the width at 512 px relative to StyleGAN3's 2^15 base
        base = int(2 ** (15 - int(math.log2(512) - math.log2(image_size))) * (channels / 64))
        widths, sizes, rates, cutoffs, half_widths = get_layer_params(image_size, num_layers, base, max_channels, image_channels, margin_size,
                                                                      first_cutoff, first_stopband, last_stopband_rel, num_critical=2)
        self.input = SynthesisInput(style_dim, int(widths[0]), sizes[0], rates[0], cutoffs[0])
        layers = []
        for i in range(num_layers + 1):
            j, rgb = max(i - 1, 0), i == num_layers
            layers.append(StyleLayer(int(widths[j]), style_dim, int(widths[i]), 1 if rgb else kernel_size, int(sizes[j]), int(sizes[i]),
                                     rates[j], rates[i], cutoffs[j], cutoffs[i], half_widths[j], half_widths[i], rgb, i >= num_layers - 2))
        self.net = nn.ModuleList(layers)
        self.register_buffer('output_scale', torch.tensor([output_scale]))

    def forward(self, w):
        if w.ndim == 2:
            w = w.unsqueeze(1).repeat(1, self.num_ws, 1)
        ws = w.unbind(dim=1)
        x = self.input(ws[0])
        for layer, w_i in zip(self.net, ws[1:]):
            x = layer(x, w_i)
        return x * self.output_scale


class Generator(nn.Module):
    def __init__(self, image_size, latent_dim, num_layers=14, map_num_layers=2, channels=32, max_channels=512, style_dim=512,
                 pixel_norm=True, image_channels=3, output_scale=0.25, margin_size=10, first_cutoff=2, first_stopband=2 ** 2.1,
                 last_stopband_rel=2 ** 0.3, kernel_size=3) -> None:
        super().__init__()
        self.map = Mapping(latent_dim, style_dim, map_num_layers, pixel_norm)
        self.synthesis = Synthesis(image_size, num_layers, channels, max_channels, style_dim, image_channels, output_scale, margin_size,
                                   first_cutoff, first_stopband, last_stopband_rel, kernel_size)

    def forward(self, z, truncation_psi=1.):
        return self.synthesis(self.map(z, truncation_psi))
