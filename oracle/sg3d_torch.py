"""CPU ORACLE (test infrastructure, NOT a product path) -- plain-PyTorch restatement of the reference's StyleGAN3-style
discriminator and of ``conv2d_resample`` for the down-sampling cases it uses.

Only tests/ may import this.  Restated from (paths relative to the STomoya/animeface checkout):
  upfirdn2d        thirdparty/stylegan3_ops/ops/upfirdn2d.py:161-207 (_upfirdn2d_ref; up = 1 here)
  conv2d_resample  thirdparty/stylegan3_ops/ops/conv2d_resample.py:40-141 (up = 1 fast paths; every case through the generic plan)
  bias_act         thirdparty/stylegan3_ops/ops/bias_act.py:86-115 (linear / lrelu)
  discriminator    implementations/StyleGAN3/model.py:16-30 (Linear), :382-510 (ConvAct, ResBlock, MinibatchStdDev,
                   DiscEpilogue, Discriminator)
PARITY PIN: tests/golden/sg3d.npz, produced by tests/golden/make_golden.py from the reference itself; checked by
tests/test_oracle_golden.py.  Functional: a flat ``state_dict`` in, tensors out; works in any float dtype / device.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


def upfirdn2d(x, f, down=1, padding=(0, 0, 0, 0), flip_filter=False, gain=1.0):
    """up = 1: pad / crop, true convolution with f (correlation when flip_filter), keep every down-th sample."""
    px0, px1, py0, py1 = padding
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0):x.shape[2] - max(-py1, 0), max(-px0, 0):x.shape[3] - max(-px1, 0)]
    if f is not None:
        f = f.to(x.dtype) * gain
        if not flip_filter:
            f = f.flip([0, 1])
        c = x.shape[1]
        x = F.conv2d(x, f[None, None].repeat(c, 1, 1, 1), groups=c)
    return x[:, :, ::down, ::down]


def conv2d_resample(x, w, f=None, down=1, padding=0):
    kh, kw = w.shape[2], w.shape[3]
    fh, fw = (1, 1) if f is None else (f.shape[0], f.shape[1])
    if isinstance(padding, int):
        px0 = px1 = py0 = py1 = padding
    else:
        px0, px1, py0, py1 = padding
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2
    if kw == 1 and kh == 1 and down > 1:
        return F.conv2d(upfirdn2d(x, f, down, (px0, px1, py0, py1)), w)
    if down > 1:
        return F.conv2d(upfirdn2d(x, f, 1, (px0, px1, py0, py1)), w, stride=down)
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return F.conv2d(x, w, padding=(py0, px0))
    return F.conv2d(upfirdn2d(x, None, 1, (px0, px1, py0, py1)), w)


def upfirdn2d_full(x, f, up=1, down=1, padding=(0, 0, 0, 0), flip_filter=False, gain=1.0):
    """thirdparty/stylegan3_ops/ops/upfirdn2d.py:161-207 (_upfirdn2d_ref) with up-sampling: zero-insert (samples at 0, up,
    2 up, ...; up - 1 trailing zeros), pad / crop, true convolution with f * gain, keep every down-th sample."""
    n, c, h, w_ = x.shape
    if up > 1:
        z = x.new_zeros(n, c, h, up, w_, up)
        z[:, :, :, 0, :, 0] = x
        x = z.reshape(n, c, h * up, w_ * up)
    return upfirdn2d(x, f, down, padding, flip_filter, gain)


def conv2d_resample_full(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True):
    """thirdparty/stylegan3_ops/ops/conv2d_resample.py:40-141 through its generic plan (:125-129), which every fast path
    must reproduce: zero-insert + low-pass (gain up^2) with all the padding in front, unpadded convolution, decimation."""
    kh, kw = w.shape[2], w.shape[3]
    fh, fw = (1, 1) if f is None else (f.shape[0], f.shape[1])
    px0, px1, py0, py1 = (padding,) * 4 if isinstance(padding, int) else padding
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2
    x = upfirdn2d_full(x, f if up > 1 else None, up, 1, (px0, px1, py0, py1), False, up ** 2)
    x = F.conv2d(x, w if flip_weight else w.flip([2, 3]), groups=groups)
    if down > 1:
        x = upfirdn2d_full(x, f, 1, down)
    return x


def conv_transpose2d(x, w, b=None, stride=1, padding=0, output_padding=0, groups=1):
    """thirdparty/stylegan3_ops/ops/conv2d_gradfix.py:34-37 with enabled = False: the ATen op itself."""
    return F.conv_transpose2d(x, w, b, stride=stride, padding=padding, output_padding=output_padding, groups=groups)


def bias_act(x, b, act, gain):
    if b is not None:
        x = x + b.reshape([1, -1] + [1] * (x.ndim - 2))
    if act == 'lrelu':
        x = F.leaky_relu(x, 0.2)
    return x * gain


def _conv_act(sd, prefix, x, down, act, act_gain, f):
    w = sd[prefix + '.weight']
    b = sd.get(prefix + '.bias')
    k = w.shape[2]
    y = conv2d_resample(x, w * (1.0 / math.sqrt(w[0].numel())), f if down > 1 else None, down, k // 2)
    return bias_act(y, b, act, act_gain)


def mbstd(x, group_size=4, num_channels=1):
    N, C, H, W = x.shape
    G = group_size if N % group_size == 0 else N
    c = C // num_channels
    y = x.reshape(G, -1, num_channels, c, H, W)
    y = y - y.mean(dim=0)
    y = (y.square().mean(dim=0) + 1e-8).sqrt().mean(dim=[2, 3, 4])
    y = y.reshape(-1, num_channels, 1, 1).repeat(G, 1, H, W)
    return torch.cat([x, y], dim=1)


def discriminator(sd, x, mbsd_group_size=4, mbsd_channels=1):
    f = None
    for k, v in sd.items():
        if k.endswith('down_filter'):
            f = v
            break
    x = _conv_act(sd, 'from_rgb', x, 1, 'lrelu', SQRT2, f)
    i = 0
    while f'resblocks.{i}.conv1.weight' in sd:
        p = f'resblocks.{i}'
        h = _conv_act(sd, p + '.conv1', x, 1, 'lrelu', SQRT2, f)
        h = _conv_act(sd, p + '.conv2', h, 2, 'lrelu', math.sqrt(0.5), f)
        x = _conv_act(sd, p + '.skip', x, 2, 'linear', math.sqrt(0.5), f) + h
        i += 1
    x = mbstd(x, mbsd_group_size, mbsd_channels)
    x = _conv_act(sd, 'epilogue.epilogue.1', x, 1, 'lrelu', SQRT2, f)
    x = x.reshape(x.shape[0], -1)
    for idx, act in ((3, 'lrelu'), (4, 'linear')):
        w, b = sd[f'epilogue.epilogue.{idx}.weight'], sd[f'epilogue.epilogue.{idx}.bias']
        x = bias_act(F.linear(x, w * (1.0 / math.sqrt(w.shape[1]))), b, act, SQRT2 if act == 'lrelu' else 1.0)
    return x
