"""``conv2d`` with gradients of any order, behind the name the reference's networks import.

Reference: thirdparty/stylegan3_ops/ops/conv2d_gradfix.py (``conv2d`` :29, ``conv_transpose2d`` :34,
``no_weight_gradients`` :19).  There the op wraps cuDNN so that R1 / path-length double backward work; here every call
lands on the closed convolution family of ``ops.conv2d`` (forward conv, data-gradient conv, weight gradient -- each one's
backward written with the other two), so the same property holds on the libsg2b200 kernels.

Supported: dilation = 1, square odd kernels k in {1, 3}, any stride / padding / groups.  The library convolves with stride 1
and 'same' padding; everything else is expressed around that call with ops that are themselves differentiable to any
order -- a larger padding pads the input first, a smaller one crops the 'same' result, a stride keeps every stride-th
sample, groups run group by group, and the transposed convolution is zero insertion (``upfirdn2d`` with up = stride and no
filter, which also applies the padding / cropping) followed by a correlation with the transposed, tap-flipped weight.
(The strided cases pay stride^2 times their flops; they are off the BASELINE config 2 hot path.)
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from . import conv2d as C

enabled = True                      # kept for API compatibility (the reference switches its custom op on with this)
weight_gradients_disabled = False   # Forcefully disable computation of gradients with respect to the weights.


@contextlib.contextmanager
def no_weight_gradients(disable=True):
    global weight_gradients_disabled
    old = weight_gradients_disabled
    if disable:
        weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def _pair(v):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    assert len(v) == 2 and all(isinstance(e, int) for e in v)
    return v


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """torch.nn.functional.conv2d semantics (cross-correlation), reference signature conv2d_gradfix.py:29."""
    assert isinstance(input, torch.Tensor) and input.ndim == 4 and weight.ndim == 4
    if _pair(dilation) != (1, 1):
        raise NotImplementedError('conv2d_gradfix.conv2d: dilation = 1 only')
    if groups != 1:
        assert input.shape[1] % groups == 0 and weight.shape[0] % groups == 0 and weight.shape[1] * groups == input.shape[1]
        xs, ws = input.chunk(groups, dim=1), weight.chunk(groups, dim=0)
        bs = bias.chunk(groups, dim=0) if bias is not None else [None] * groups
        return torch.cat([conv2d(x_, w_, b_, stride, padding, dilation, 1) for x_, w_, b_ in zip(xs, ws, bs)], dim=1)
    co, ci, kh, kw = weight.shape
    if kh != kw or kh not in (1, 3):
        raise NotImplementedError('conv2d_gradfix.conv2d: square kernels of size 1 or 3 only')
    sy, sx = _pair(stride)
    py, px = _pair(padding)
    assert sy >= 1 and sx >= 1 and py >= 0 and px >= 0
    if weight_gradients_disabled:
        weight = weight.detach()
    half = kh // 2
    ey, ex = py - half, px - half                    # padding relative to 'same'
    if ey > 0 or ex > 0:
        input = F.pad(input, [max(ex, 0), max(ex, 0), max(ey, 0), max(ey, 0)])
    y = C.conv2d(input, weight, 1.0)
    cy, cx = max(-ey, 0), max(-ex, 0)
    if cy or cx or sy > 1 or sx > 1:
        y = y[:, :, cy:y.shape[2] - cy:sy, cx:y.shape[3] - cx:sx]
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1)
    return y


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """torch.nn.functional.conv_transpose2d semantics, reference signature conv2d_gradfix.py:34.
    input [N, Ci, H, W], weight [Ci, Co / groups, k, k]; out size (H - 1) * stride - 2 * padding + k + output_padding."""
    from . import upfirdn2d as U
    assert isinstance(input, torch.Tensor) and input.ndim == 4 and weight.ndim == 4
    if _pair(dilation) != (1, 1):
        raise NotImplementedError('conv2d_gradfix.conv_transpose2d: dilation = 1 only')
    if groups != 1:
        assert input.shape[1] % groups == 0 and weight.shape[0] == input.shape[1]
        xs, ws = input.chunk(groups, dim=1), weight.chunk(groups, dim=0)
        y = torch.cat([conv_transpose2d(x_, w_, None, stride, padding, output_padding, 1, dilation) for x_, w_ in zip(xs, ws)], dim=1)
        return y if bias is None else y + bias.reshape(1, -1, 1, 1)
    ci, co, kh, kw = weight.shape
    if kh != kw or kh not in (1, 3):
        raise NotImplementedError('conv2d_gradfix.conv_transpose2d: square kernels of size 1 or 3 only')
    sy, sx = _pair(stride)
    py, px = _pair(padding)
    oy, ox = _pair(output_padding)
    assert sy >= 1 and sx >= 1 and py >= 0 and px >= 0 and 0 <= oy < max(sy, 2) and 0 <= ox < max(sx, 2)
    if weight_gradients_disabled:
        weight = weight.detach()
    # zero insertion leaves the samples at 0, s, 2s, ... of a length H*s signal (s - 1 trailing zeros); the full correlation
    # needs k - 1 - p zeros in front and k - 1 - p + output_padding behind the LAST sample: negative amounts crop
    fy, fx = kh - 1 - py, kw - 1 - px
    pads = [fx, fx - (sx - 1) + ox, fy, fy - (sy - 1) + oy]
    if sy > 1 or sx > 1 or any(pads):
        input = U.upfirdn2d(input, None, up=[sx, sy], padding=pads)
    y = conv2d(input, weight.transpose(0, 1).flip([2, 3]), None, 1, 0)
    return y if bias is None else y + bias.reshape(1, -1, 1, 1)
