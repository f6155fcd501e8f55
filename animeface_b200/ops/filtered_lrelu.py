"""``filtered_lrelu`` behind the reference's name and signature (thirdparty/stylegan3_ops/ops/filtered_lrelu.py:50-268).

bias -> zero-insert up-sampling by `up` + padding + FIR `fu` (gain up^2) -> leaky ReLU * gain, clamp -> FIR `fd` +
decimation by `down`: the alias-suppressed non-linearity of the StyleGAN3 generator (implementations/StyleGAN3/model.py:186-190).

Execution: the four stages run on the library's own ``bias_act`` and ``upfirdn2d`` kernels -- the arithmetic of the reference's
``_filtered_lrelu_ref`` (:121-147), which its fused CUDA kernel (filtered_lrelu.cu:133-1093) reproduces -- so gradients of any
order exist through the ops' own closed families (the reference builds a dedicated backward graph with bit-packed signs for
that, :150-268).  The up-sampled intermediate is materialised (up^2 x the input); fusing the four stages per shared-memory
tile is the remaining optimisation for BASELINE config 5 and is documented in DESIGN.md as not built.
"""
from __future__ import annotations

import numpy as np
import torch

from . import bias_act as _ba
from . import upfirdn2d as _up
from .upfirdn2d import _quad, _taps


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, impl='cuda'):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert impl in ('ref', 'cuda')
    if impl == 'ref':
        raise RuntimeError("filtered_lrelu: impl='ref' is not a product path here; see oracle/sg3g_torch.py")
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    assert gain == float(gain) and gain > 0
    assert slope == float(slope) and slope >= 0
    assert clamp is None or (clamp == float(clamp) and clamp >= 0)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.dtype == x.dtype and tuple(b.shape) == (x.shape[1],)
    px0, px1, py0, py1 = _quad(padding)
    fu_w, fu_h = _taps(fu)
    fd_w, fd_h = _taps(fd)
    n, c, in_h, in_w = x.shape
    out_w = (in_w * up + (px0 + px1) - (fu_w - 1) - (fd_w - 1) + (down - 1)) // down
    out_h = (in_h * up + (py0 + py1) - (fu_h - 1) - (fd_h - 1) + (down - 1)) // down
    y = _ba.bias_act(x=x, b=b) if b is not None else x
    y = _up.upfirdn2d(x=y, f=fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    y = _ba.bias_act(x=y, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    y = _up.upfirdn2d(x=y, f=fd, down=down, flip_filter=flip_filter)
    assert tuple(y.shape) == (n, c, out_h, out_w), (tuple(y.shape), (n, c, out_h, out_w))
    return y
