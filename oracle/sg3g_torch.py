"""CPU ORACLE (test infrastructure, NOT a product path) -- plain-PyTorch restatement of ``filtered_lrelu`` and of the
StyleGAN3 generator's layers.

Only tests/ may import this.  Restated from (paths relative to the STomoya/animeface checkout):
  filtered_lrelu   thirdparty/stylegan3_ops/ops/filtered_lrelu.py:121-147 (_filtered_lrelu_ref: bias, up-FIR with gain up^2,
                   leaky ReLU * gain + clamp, down-FIR) on oracle/sg3d_torch.py's upfirdn2d restatement
  modulated conv   implementations/StyleGAN3/model.py:32-72 (per-sample weights, demodulation eps 1e-8, input gain after it)
PARITY PIN: tests/golden/sg3g.npz, produced by tests/golden/make_golden.py from the reference itself; checked by
tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .sg3d_torch import upfirdn2d_full


def _as2d(f):
    return None if f is None else (torch.outer(f, f) if f.ndim == 1 else f)


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=2 ** 0.5, slope=0.2, clamp=None, flip_filter=False):
    pad = (padding,) * 4 if isinstance(padding, int) else tuple(padding)
    if b is not None:
        x = x + b.reshape(1, -1, 1, 1)
    x = upfirdn2d_full(x, _as2d(fu), up, 1, pad, flip_filter, up ** 2)
    x = F.leaky_relu(x, slope) * gain
    if clamp is not None:
        x = x.clamp(-clamp, clamp)
    return upfirdn2d_full(x, _as2d(fd), 1, down, (0, 0, 0, 0), flip_filter, 1.0)


def modulated_conv(x, w, s, padding, demod=True, input_gain=None):
    B = x.shape[0]
    co, ci, k, _ = w.shape
    wb = w[None] * s[:, None, :, None, None] * (1.0 / (ci * k * k) ** 0.5)
    if demod:
        wb = wb * wb.square().sum([2, 3, 4]).add(1e-8).rsqrt()[:, :, None, None, None]
    if input_gain is not None:
        wb = wb * input_gain.expand(B, ci)[:, None, :, None, None]
    y = F.conv2d(x.reshape(1, B * ci, *x.shape[2:]), wb.reshape(B * co, ci, k, k), padding=padding, groups=B)
    return y.reshape(B, co, *y.shape[2:])
