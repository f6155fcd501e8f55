"""``conv2d`` with gradients of any order, behind the name the reference's networks import.

Reference: thirdparty/stylegan3_ops/ops/conv2d_gradfix.py (``conv2d`` :29, ``conv_transpose2d`` :34,
``no_weight_gradients`` :19).  There the op wraps cuDNN so that R1 / path-length double backward work; here every call
lands on the closed convolution family of ``ops.conv2d`` (forward conv, data-gradient conv, weight gradient -- each one's
backward written with the other two), so the same property holds on the libsg2b200 kernels.

Supported: groups = 1, dilation = 1, square odd kernels k in {1, 3}.  The library convolves with stride 1 and 'same'
padding; other stride / padding combinations are expressed around it -- a larger padding pads the input first, a smaller
one crops the 'same' result, a stride keeps every stride-th sample.  (A strided launch is round-2 work: today the
stride-2 convolutions of the StyleGAN3-style discriminator pay 4x their flops.)  ``conv_transpose2d`` is only used by the
up-sampling branch of ``conv2d_resample`` (StyleGAN3 generator) and is not built.
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
    if groups != 1 or _pair(dilation) != (1, 1):
        raise NotImplementedError('conv2d_gradfix.conv2d: groups = 1 and dilation = 1 only')
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
    raise NotImplementedError('conv2d_gradfix.conv_transpose2d (the up-sampling branch of conv2d_resample) is not built: '
                              'it belongs to the StyleGAN3 generator, SURVEY 8f n3')
