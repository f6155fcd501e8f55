"""2-D convolution with optional down-sampling, behind the reference's name and signature.

Reference: thirdparty/stylegan3_ops/ops/conv2d_resample.py:40-137.  Padding is applied once, in front; the low-pass
filter runs through ``upfirdn2d`` (the register-ring fast path for the [1,3,3,1] x [1,3,3,1] blur of the StyleGAN3-style
discriminator, csrc/upfirdn2d.cu), the convolution through ``conv2d_gradfix``.  The up-sampling branches (:110-127) need
``conv_transpose2d`` and belong to the StyleGAN3 generator: they raise NotImplementedError.
"""
from __future__ import annotations

import torch

from . import conv2d_gradfix, upfirdn2d
from .upfirdn2d import _quad, _taps


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d_resample.py:24-36: conv2d() correlates (flip_weight=True); flip the taps for a true convolution."""
    kh, kw = w.shape[2], w.shape[3]
    if not flip_weight and (kw > 1 or kh > 1):
        w = w.flip([2, 3])
    if transpose:
        return conv2d_gradfix.conv_transpose2d(x, w, stride=stride, padding=padding, groups=groups)
    return conv2d_gradfix.conv2d(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in (1, 2) and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1
    assert isinstance(down, int) and down >= 1
    assert isinstance(groups, int) and groups >= 1
    kh, kw = int(w.shape[2]), int(w.shape[3])
    fw, fh = _taps(f)
    px0, px1, py0, py1 = _quad(padding)
    if up > 1:
        raise NotImplementedError('conv2d_resample: up > 1 needs conv_transpose2d (StyleGAN3 generator, SURVEY 8f n3)')
    # adjust padding to account for down-sampling (:75-79)
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    # 1x1 convolution with down-sampling only => down-sample first, then convolve (:82-85)
    if kw == 1 and kh == 1 and down > 1:
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
    # down-sampling only => low-pass, then strided convolution (:94-97)
    if down > 1:
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, stride=down, groups=groups, flip_weight=flip_weight)
    # no resampling, symmetric non-negative padding => plain conv2d (:130-132)
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return _conv2d_wrapper(x=x, w=w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    # generic path (:135-139): pad / crop with an identity upfirdn2d, then convolve without padding
    x = upfirdn2d.upfirdn2d(x=x, f=None, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
    return _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
