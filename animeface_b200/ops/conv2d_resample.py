"""2-D convolution with optional up- / down-sampling, behind the reference's name and signature.

Reference: thirdparty/stylegan3_ops/ops/conv2d_resample.py:40-141.  Padding is applied once, in front; low-pass filters run
through ``upfirdn2d`` (the register-ring fast path for the [1,3,3,1] x [1,3,3,1] blur of the StyleGAN3-style discriminator,
csrc/upfirdn2d.cu), convolutions through ``conv2d_gradfix`` -- so every branch is differentiable to any order on the
libsg2b200 kernels.  The same five execution plans as the reference, chosen by (kernel size, up, down):
  pointwise + down   : filter & decimate, then the 1x1 convolution on the small image          (:82-85)
  pointwise + up     : the 1x1 convolution on the small image, then zero-insert & filter       (:88-91)
  down only          : low-pass at full resolution, then a stride-`down` convolution           (:94-97)
  up [+ down]        : transposed stride-`up` convolution, low-pass (gain up^2) [, decimate]   (:100-117)
  plain              : one convolution when the padding is symmetric and non-negative          (:120-122)
  anything else      : zero-insert / pad with upfirdn2d, convolve unpadded [, decimate]        (:125-129)
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
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def _to_transposed_layout(w, groups):
    """[Co, Ci/g, kh, kw] -> the conv_transpose2d layout [Ci, Co/g, kh, kw] (per group)."""
    if groups == 1:
        return w.transpose(0, 1)
    co, cig, kh, kw = w.shape
    return w.reshape(groups, co // groups, cig, kh, kw).transpose(1, 2).reshape(groups * cig, co // groups, kh, kw)


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
    # centre the filter footprint on the resampled grid (:75-84)
    if up > 1:
        px0, px1 = px0 + (fw + up - 1) // 2, px1 + (fw - up) // 2
        py0, py1 = py0 + (fh + up - 1) // 2, py1 + (fh - up) // 2
    if down > 1:
        px0, px1 = px0 + (fw - down + 1) // 2, px1 + (fw - down) // 2
        py0, py1 = py0 + (fh - down + 1) // 2, py1 + (fh - down) // 2
    fir = lambda t, **kw_: upfirdn2d.upfirdn2d(x=t, f=f, flip_filter=flip_filter, **kw_)
    conv = lambda t, ww=w, **kw_: _conv2d_wrapper(x=t, w=ww, groups=groups, flip_weight=flip_weight, **kw_)
    pointwise = kh == 1 and kw == 1

    if pointwise and down > 1 and up == 1:
        return conv(fir(x, down=down, padding=[px0, px1, py0, py1]))
    if pointwise and up > 1 and down == 1:
        return fir(conv(x), up=up, padding=[px0, px1, py0, py1], gain=up ** 2)
    if down > 1 and up == 1:
        return conv(fir(x, padding=[px0, px1, py0, py1]), stride=down)
    if up > 1:
        # the transposed convolution produces (H - 1) * up + k samples: the part of the padding it can absorb itself is
        # passed to it, the rest (possibly negative = cropping) goes to the low-pass that follows
        px0, px1, py0, py1 = px0 - (kw - 1), px1 - (kw - up), py0 - (kh - 1), py1 - (kh - up)
        pxt, pyt = max(min(-px0, -px1), 0), max(min(-py0, -py1), 0)
        y = _conv2d_wrapper(x=x, w=_to_transposed_layout(w, groups), stride=up, padding=[pyt, pxt], groups=groups, transpose=True,
                            flip_weight=(not flip_weight))
        y = fir(y, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2)
        return fir(y, down=down) if down > 1 else y
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return conv(x, padding=[py0, px0])
    # generic: pad / crop with an identity upfirdn2d, convolve without padding
    return conv(upfirdn2d.upfirdn2d(x=x, f=None, padding=[px0, px1, py0, py1], flip_filter=flip_filter))
