"""StyleGAN2 generator / discriminator of the reference, re-built on the libsg2b200 kernels.

Drop-in for implementations/StyleGAN2/model.py: same class names, constructor arguments, forward
signatures and -- because the module tree is mirrored attribute for attribute -- the same 85/44-key
``state_dict`` (reference checkpoints load here and vice versa; that is also how parity is tested).

What changes is the execution:
  * activations live in channels_last (NHWC) memory so the convolution kernels see K-contiguous tiles;
  * ``Upsample2x -> Blur2d`` is one fused kernel (ops.resample), ``conv -> bias -> noise -> lrelu`` is one
    conv kernel with a fused epilogue, ``AvgPool -> add -> /sqrt2`` is one kernel, minibatch-stddev two;
  * the modulated convolution never builds the per-sample weight tensor [B,Co,Ci,k,k]
    (reference model.py:115-120): the style scales the activation tile, demodulation is an epilogue scale.
"""
from __future__ import annotations

import functools
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import rng
from .ops import conv2d as C
from .ops.bias_act import bias_act
from .ops.linear import linear_bias_act, pixel_norm
from .ops.mbstd import minibatch_stddev
from .ops.resample import avgpool2, upsample2x_bilinear, upsample2x_blur

SLOPE = 0.2


class ELR(nn.Module):
    """Equalised learning rate wrapper (reference model.py:29-37): y = layer(x * coef).

    Only parameter storage and the constant live here; convolutions are executed by the caller through
    the fused ops with ``coef`` folded into the packed weight."""

    def __init__(self, layer, gain=1.):
        super().__init__()
        self.coef = gain / (layer.weight[0].numel() ** 0.5)
        self.layer = layer

    def forward(self, x):
        if isinstance(self.layer, nn.Linear):
            return linear_bias_act(x, self.layer.weight, self.layer.bias, self.coef)
        pad = self.layer.padding[0]
        assert self.layer.stride == (1, 1) and 2 * pad + 1 == self.layer.kernel_size[0], 'stride-1 same conv only'
        return C.conv2d_bias_act(x, self.layer.weight, self.layer.bias, self.coef, None)


def Linear(name, *args, **kwargs):
    layer = nn.Linear(*args, **kwargs)
    return ELR(layer) if name == 'elr' else layer


def Conv2d(name, *args, **kwargs):
    layer = nn.Conv2d(*args, **kwargs)
    return ELR(layer) if name == 'elr' else layer


class Upsample2x(nn.Module):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (reference model.py:56-58)."""

    def __init__(self, name='bilinear'):
        super().__init__()
        if name != 'bilinear':
            raise NotImplementedError("only the 'bilinear' upsampler of the reference path is built")

    def forward(self, x):
        return upsample2x_bilinear(x)


class Downsample2x(nn.Module):
    def __init__(self, name='avg'):
        super().__init__()
        if name != 'avg':
            raise NotImplementedError("only the 'avg' downsampler of the reference path is built")

    def forward(self, x):
        return avgpool2(x)


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.size(0), -1)


class MapLinear(nn.Module):
    """(x W coef + b) * lr (reference model.py:71-78)."""

    def __init__(self, *args, lr=0.01, **kwargs):
        super().__init__()
        self.linear = Linear('elr', *args, **kwargs)
        self.lr = lr

    def forward(self, x, slope=None):
        # (layer(x * coef)) * lr [-> LeakyReLU]: one launch (ops/linear.py)
        lin = self.linear
        return linear_bias_act(x, lin.layer.weight, lin.layer.bias, lin.coef, self.lr, slope)


class supplied_noise(rng.replay):
    """Context manager for tests: InjectNoise takes its tensors from ``seq`` instead of torch.randn."""


class InjectNoise(nn.Module):
    """x + randn(B,1,H,W) (reference model.py:81-88; ``scale`` exists but is never applied there either)."""

    def __init__(self):
        super().__init__()
        self.scale = nn.Parameter(torch.zeros(1))

    def sample(self, B, H, W, device):
        return rng.randn(B, 1, H, W, device=device)

    def forward(self, x):
        B, _, H, W = x.size()
        return x + self.sample(B, H, W, x.device)


class ModulatedConv2d(nn.Module):
    """Modulated / demodulated convolution (reference model.py:91-135), style -> affine -> scale."""

    def __init__(self, in_channels, out_channels, style_dim, kernel_size, stride=1, demod=True, gain=1.):
        super().__init__()
        if stride != 1:
            raise NotImplementedError('stride-1 only (the reference never uses another value)')
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.demod = demod
        self.affine = Linear('elr', style_dim, in_channels)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(1, out_channels, 1, 1))
        self.coef = gain / (self.weight[0].numel() ** 0.5)
        assert gain == 1., 'gain != 1 is not used by the reference path'

    def forward(self, x, y, noise=None, slope=None):
        s = self.affine(y) + 1
        return C.modulated_conv2d(x, self.weight, s, self.bias, noise, self.demod, slope,
                                  out_nchw=(self.out_channels % 4 != 0))


class Blur2d(nn.Module):
    """[1,2,1]x[1,2,1]/16 depthwise blur (reference model.py:138-149).  Inside StyleBlock it is fused with
    the preceding upsample; stand-alone it runs as an upfirdn2d filter2d."""

    def __init__(self):
        super().__init__()
        k = torch.tensor([[[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]])
        self.register_buffer('kernel', k / k.sum())

    def forward(self, x):
        from .ops.upfirdn2d import filter2d
        return filter2d(x, self.kernel[0].to(torch.float32))


class StyleBlock(nn.Module):
    """upsample -> blur -> [modconv -> noise -> lrelu] * num_conv (reference model.py:154-180)."""

    def __init__(self, in_channels, out_channels, style_dim, num_conv=2, up_name='bilinear'):
        super().__init__()
        mods = [Upsample2x(up_name), Blur2d()]
        for i in range(num_conv):
            mods += [ModulatedConv2d(in_channels if i == 0 else out_channels, out_channels, style_dim, 3),
                     InjectNoise(), nn.LeakyReLU(SLOPE, inplace=True)]
        self.block = nn.ModuleList(mods)

    def forward(self, x, y):
        x = upsample2x_blur(x)                           # block[0] + block[1] in one pass
        for i in range(2, len(self.block), 3):
            conv, inject = self.block[i], self.block[i + 1]
            B, _, H, W = x.shape
            x = conv(x, y, noise=inject.sample(B, H, W, x.device), slope=SLOPE)
        return x


class DBlock(nn.Module):
    """Residual discriminator block (reference model.py:186-212)."""

    def __init__(self, in_channels, out_channels, num_conv=2, down_name='avg'):
        super().__init__()
        layers = []
        for i in range(num_conv):
            layers += [Conv2d('elr', in_channels if i == 0 else out_channels, out_channels, 3, padding=1),
                       nn.LeakyReLU(SLOPE, inplace=True)]
        self.block = nn.Sequential(*layers)
        self.down = Downsample2x(down_name)
        self.skip = Conv2d('elr', in_channels, out_channels, 1)

    def forward(self, x):
        if len(self.block) == 4 and isinstance(self.down, Downsample2x) and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            # the standard block (two convolutions) as one autograd node: ops.conv2d.DBlockFn
            c1, c2, sk = self.block[0], self.block[2], self.skip
            return C.dblock(x, c1.layer.weight, c1.layer.bias, c2.layer.weight, c2.layer.bias, sk.layer.weight, sk.layer.bias,
                            c1.coef, c2.coef, sk.coef, SLOPE, 1.0 / math.sqrt(2.0))
        t = C.conv2d_bias_act(x, self.skip.layer.weight, self.skip.layer.bias, self.skip.coef, None)
        for i in range(0, len(self.block), 2):
            conv = self.block[i]
            x = C.conv2d_bias_act(x, conv.layer.weight, conv.layer.bias, conv.coef, SLOPE)
        return avgpool2(x, t, 1.0 / math.sqrt(2.0))      # (down(x) + down(t)) / sqrt(2), one kernel


_mbstd_chunks = 1


class independent_batches:
    """Context manager: the batch handed to the Discriminator is ``chunks`` independent mini-batches concatenated
    (the training step runs D(real) and D(fake) of implementations/StyleGAN2/utils.py:65,69 as ONE call).  Every layer of D
    is per-sample except MiniBatchStdDev, whose group statistics must not mix the mini-batches: inside this context it
    evaluates each chunk on its own, exactly as the separate calls would."""

    def __init__(self, chunks):
        self.chunks = int(chunks)

    def __enter__(self):
        global _mbstd_chunks
        self._prev, _mbstd_chunks = _mbstd_chunks, self.chunks
        return self

    def __exit__(self, *exc):
        global _mbstd_chunks
        _mbstd_chunks = self._prev


class MiniBatchStdDev(nn.Module):
    def __init__(self, group_size, eps=1e-4):
        super().__init__()
        self.group_size = group_size
        self.eps = eps

    def forward(self, x):
        if _mbstd_chunks > 1:
            assert x.shape[0] % _mbstd_chunks == 0
            return torch.cat([minibatch_stddev(c, self.group_size, self.eps) for c in x.chunk(_mbstd_chunks, dim=0)], dim=0)
        return minibatch_stddev(x, self.group_size, self.eps)


class ToImage(nn.Module):
    """1x1 modulated conv to RGB, skip-sum, optional bilinear x2 (reference model.py:239-250)."""

    def __init__(self, in_channels, image_channels, style_dim, upsample=True, up_name='bilinear'):
        super().__init__()
        self.conv = ModulatedConv2d(in_channels, image_channels, style_dim, 1, demod=False)
        self.upsample = Upsample2x(up_name) if upsample else None

    def forward(self, x, y, pre=None):
        x = self.conv(x, y)
        if pre is not None:
            x = x + pre
        if self.upsample is not None:
            x = self.upsample(x)
        return x


class PixelNorm(nn.Module):
    def forward(self, x):
        return pixel_norm(x, 1e-4)


class Mapping(nn.Module):
    def __init__(self, style_dim, num_layers=8, normalize=True, lr=0.01):
        super().__init__()
        self.normalize = PixelNorm() if normalize else None
        layers = []
        for _ in range(num_layers):
            layers += [MapLinear(style_dim, style_dim, lr=lr), nn.LeakyReLU(SLOPE, inplace=True)]
        self.map = nn.Sequential(*layers)

    def forward(self, x):
        if self.normalize is not None:
            x = self.normalize(x)
        mods = list(self.map)
        i = 0
        while i < len(mods):
            fuse = isinstance(mods[i], MapLinear) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU)
            if fuse:
                x = mods[i](x, slope=mods[i + 1].negative_slope)      # linear * lr -> lrelu in one kernel
                i += 2
            else:
                x = mods[i](x)
                i += 1
        return x


class Synthesis(nn.Module):
    def __init__(self, image_size, image_channels, style_dim, channels=32, max_channels=512, num_conv=2):
        super().__init__()
        cap = functools.partial(min, max_channels)
        channels = channels * (2 ** int(np.log2(image_size) - 2))
        och = cap(channels)
        self.input = ModulatedConv2d(style_dim, och, style_dim, 3)
        self.input_to_image = ToImage(och, image_channels, style_dim)
        self.num_layers = 1
        self.blocks = nn.ModuleList()
        self.to_images = nn.ModuleList()
        resl = 4
        while resl < image_size:
            resl *= 2
            channels //= 2
            ich, och = och, cap(channels)
            self.blocks.append(StyleBlock(ich, och, style_dim, num_conv))
            self.to_images.append(ToImage(och, image_channels, style_dim, upsample=resl < image_size))
            self.num_layers += 1
        self.tanh = nn.Tanh()

    def forward(self, x, y, injection=None):
        if isinstance(y, (list, tuple)):                 # style mixing (reference model.py:315-322)
            assert len(y) == 2
            if injection is None or injection > self.num_layers:
                injection = np.random.randint(0, self.num_layers)
            y = [y[0]] * injection + [y[1]] * (self.num_layers - injection)
        else:
            y = [y] * self.num_layers
        x = self.input(x, y[0])
        image = pre = self.input_to_image(x, y[0])
        for block, to_image, w in zip(self.blocks, self.to_images, y[1:]):
            x = block(x, w)
            image = pre = to_image(x, w, pre)
        return self.tanh(image)


class Generator(nn.Module):
    def __init__(self, image_size=128, image_channels=3, style_dim=512, channels=32, max_channels=512,
                 block_num_conv=2, map_num_layers=8, normalize_latent=True, map_lr=0.01):
        super().__init__()
        self.map = Mapping(style_dim, map_num_layers, normalize_latent, map_lr)
        self.synthesis = Synthesis(image_size, image_channels, style_dim, channels, max_channels, block_num_conv)
        self.const = nn.Parameter(torch.empty(1, style_dim, 4, 4))
        self.const.data.normal_(0, 1)

    def forward(self, z, injection=None):
        if isinstance(z, (list, tuple)):
            style = [self.map(z[0]), self.map(z[1])]
            B = z[0].size(0)
        else:
            style = self.map(z)
            B = z.size(0)
        x = self.const.expand(B, -1, -1, -1)
        return self.synthesis(x, style, injection), style

    def init_weight(self, map_init_func, syn_init_func):
        self.map.apply(map_init_func)
        self.synthesis.apply(syn_init_func)


class Discriminator(nn.Module):
    def __init__(self, image_size=128, image_channels=3, channels=32, max_channels=512, block_num_conv=2, mbsd_groups=4):
        super().__init__()
        cap = functools.partial(min, max_channels)
        och = channels
        self.from_rgb = nn.Sequential(Conv2d('elr', image_channels, och, 1), nn.LeakyReLU(SLOPE, inplace=True))
        resl = image_size
        blocks = []
        while resl > 4:
            resl //= 2
            channels *= 2
            ich, och = och, cap(channels)
            blocks.append(DBlock(ich, och, block_num_conv))
        blocks.append(MiniBatchStdDev(mbsd_groups))
        blocks += [Conv2d('elr', och + 1, och, 3, padding=1), nn.LeakyReLU(SLOPE, inplace=True), Flatten(),
                   Linear('elr', och * (resl ** 2), och), nn.LeakyReLU(SLOPE, inplace=True), Linear('elr', och, 1)]
        self.blocks = nn.Sequential(*blocks)

    def forward(self, x):
        rgb = self.from_rgb[0]
        x = C.conv2d_bias_act(x, rgb.layer.weight, rgb.layer.bias, rgb.coef, SLOPE)
        mods = list(self.blocks)
        i = 0
        while i < len(mods):
            m = mods[i]
            fuse = i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU)
            if (isinstance(m, MiniBatchStdDev) and i + 2 < len(mods) and isinstance(mods[i + 1], ELR)
                    and isinstance(mods[i + 1].layer, nn.Conv2d) and isinstance(mods[i + 2], nn.LeakyReLU) and x.shape[1] % 32 == 0):
                # conv over [x, stddev] (C+1 = 513 channels) = conv(x, W[:, :C]) + conv(stddev map, W[:, C:]): the first
                # term runs on the tensor cores (C is a multiple of 32), the one-channel term is a trivial fp32 conv.
                conv = mods[i + 1]
                c_in = x.shape[1]
                y = m(x)
                w = conv.layer.weight
                z = C.conv2d(x, w[:, :c_in].contiguous(), conv.coef) + C.conv2d(y[:, c_in:], w[:, c_in:].contiguous(), conv.coef)
                x = bias_act(z, conv.layer.bias, act='lrelu', alpha=SLOPE, gain=1.0)
                i += 3
                continue
            if isinstance(m, ELR) and isinstance(m.layer, nn.Conv2d):
                x = C.conv2d_bias_act(x, m.layer.weight, m.layer.bias, m.coef, SLOPE if fuse else None)
                i += 2 if fuse else 1
            elif isinstance(m, ELR):
                x = linear_bias_act(x, m.layer.weight, m.layer.bias, m.coef, 1.0, SLOPE if fuse else None)
                i += 2 if fuse else 1
            else:
                x = m(x)
                i += 1
        return x


def init_weight_N01(m, lr=1):
    """weights ~ N(0, 1/lr), biases 0 (reference model.py:404-408)."""
    if isinstance(m, (nn.Linear, nn.Conv2d, ModulatedConv2d)):
        m.weight.data.normal_(0., 1 / lr)
        m.bias.data.fill_(0.)
