"""DiffAugment ('color', 'translation', 'cutout') with the reference's signature and draw order.

Reference: thirdparty/diffaugment/DiffAugment.py:10-77.  Same arithmetic and the same sequence of random
draws (3 x rand(B,1,1,1) for colour, 2 x randint for translation / cutout).  On the training path (CUDA fp32 NCHW
images, policy a subset of 'color,translation,cutout' in that order) the whole policy is ONE fused kernel pass forward
and one backward (csrc/diffaug.cu) instead of the reference's ~10 elementwise / reduction / gather passes; the op is
affine in x, its backward kernel is the transpose and the pair is closed under differentiation.  Any other policy
order / layout takes the per-op functions below (gather-based translation), which stay differentiable w.r.t. x.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib, rng
from ._lib import amp_bwd, amp_fwd

_CANON = ('color', 'translation', 'cutout')


class _Draws:
    """The random draws of one DiffAugment call, in the reference's order (DiffAugment.py:24,30,36,42-43,58-59)."""

    def __init__(self, x, parts, cutout_ratio=0.5, translation_ratio=0.125):
        B, C, H, W = x.shape
        dev = x.device
        self.rb = self.rs = self.rc = self.ty = self.tx = self.cy = self.cx = None
        self.cut_h = self.cut_w = 0
        if 'color' in parts:
            self.rb = rng.rand(B, 1, 1, 1, device=dev).reshape(B).contiguous()
            self.rs = rng.rand(B, 1, 1, 1, device=dev).reshape(B).contiguous()
            self.rc = rng.rand(B, 1, 1, 1, device=dev).reshape(B).contiguous()
        if 'translation' in parts:
            sh, sw = int(H * translation_ratio + 0.5), int(W * translation_ratio + 0.5)
            self.ty = rng.randint(-sh, sh + 1, (B, 1, 1), device=dev).reshape(B).to(torch.int64).contiguous()
            self.tx = rng.randint(-sw, sw + 1, (B, 1, 1), device=dev).reshape(B).to(torch.int64).contiguous()
        if 'cutout' in parts:
            self.cut_h, self.cut_w = int(H * cutout_ratio + 0.5), int(W * cutout_ratio + 0.5)
            self.cy = rng.randint(0, H + (1 - self.cut_h % 2), (B, 1, 1), device=dev).reshape(B).to(torch.int64).contiguous()
            self.cx = rng.randint(0, W + (1 - self.cut_w % 2), (B, 1, 1), device=dev).reshape(B).to(torch.int64).contiguous()


def _launch(x, d, backward, linear_only):
    lib = _lib.load()
    B, C, H, W = x.shape
    x = x.contiguous()
    y = torch.empty_like(x)
    ws = torch.empty(max(int(lib.sg2_diffaugment_workspace(B, H, W)), 16), dtype=torch.uint8, device=x.device)
    _lib.check(lib.sg2_diffaugment(x.data_ptr(), y.data_ptr(), _lib.ptr(d.rb), _lib.ptr(d.rs), _lib.ptr(d.rc), _lib.ptr(d.ty), _lib.ptr(d.tx),
                                   _lib.ptr(d.cy), _lib.ptr(d.cx), d.cut_h, d.cut_w, B, C, H, W, 1 if backward else 0,
                                   1 if linear_only else 0, ws.data_ptr(), _lib.stream_ptr(x)), 'sg2_diffaugment')
    return y


class _AugFn(torch.autograd.Function):
    """y = A x + c (the fused policy); linear_only drops c (that is the derivative of the backward op below)."""

    @staticmethod
    @amp_fwd
    def forward(ctx, x, d, linear_only):
        ctx.d = d
        return _launch(x, d, False, linear_only)

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        return _AugTFn.apply(gy, ctx.d), None, None


class _AugTFn(torch.autograd.Function):
    """gx = A^T gy"""

    @staticmethod
    @amp_fwd
    def forward(ctx, gy, d):
        ctx.d = d
        return _launch(gy, d, True, False)

    @staticmethod
    @amp_bwd
    def backward(ctx, ggx):
        return _AugFn.apply(ggx, ctx.d, True), None


def _fused_ok(x, parts):
    return (x.is_cuda and x.dtype == torch.float32 and x.ndim == 4 and x.shape[1] <= 8 and len(parts) > 0
            and all(p in _CANON for p in parts) and [p for p in _CANON if p in parts] == parts)


def DiffAugment(x, policy='', channels_first=True):
    if policy and channels_first and _fused_ok(x, policy.split(',')):
        return _AugFn.apply(x, _Draws(x, policy.split(',')), False)
    if policy:
        if not channels_first:
            x = x.permute(0, 3, 1, 2)
        for p in policy.split(','):
            for f in AUGMENT_FNS[p]:
                x = f(x)
        if not channels_first:
            x = x.permute(0, 2, 3, 1)
        x = x.contiguous()
    return x


def rand_brightness(x):
    return x + (rng.rand(x.size(0), 1, 1, 1, device=x.device, dtype=x.dtype) - 0.5)


def rand_saturation(x):
    m = x.mean(dim=1, keepdim=True)
    return (x - m) * (rng.rand(x.size(0), 1, 1, 1, device=x.device, dtype=x.dtype) * 2) + m


def rand_contrast(x):
    m = x.mean(dim=[1, 2, 3], keepdim=True)
    return (x - m) * (rng.rand(x.size(0), 1, 1, 1, device=x.device, dtype=x.dtype) + 0.5) + m


def rand_translation(x, ratio=0.125):
    B, C, H, W = x.shape
    sh, sw = int(H * ratio + 0.5), int(W * ratio + 0.5)
    th = rng.randint(-sh, sh + 1, (B, 1, 1), device=x.device).reshape(B)
    tw = rng.randint(-sw, sw + 1, (B, 1, 1), device=x.device).reshape(B)
    rows = (torch.arange(H, device=x.device)[None, :] + th[:, None] + 1).clamp_(0, H + 1)     # [B,H] into padded
    cols = (torch.arange(W, device=x.device)[None, :] + tw[:, None] + 1).clamp_(0, W + 1)     # [B,W]
    xp = F.pad(x, [1, 1, 1, 1])
    xp = torch.gather(xp, 2, rows[:, None, :, None].expand(B, C, H, W + 2))
    return torch.gather(xp, 3, cols[:, None, None, :].expand(B, C, H, W))


def rand_cutout(x, ratio=0.5):
    B, C, H, W = x.shape
    ch, cw = int(H * ratio + 0.5), int(W * ratio + 0.5)
    oy = rng.randint(0, H + (1 - ch % 2), (B, 1, 1), device=x.device).reshape(B, 1)
    ox = rng.randint(0, W + (1 - cw % 2), (B, 1, 1), device=x.device).reshape(B, 1)
    ys = torch.arange(H, device=x.device)[None, :]
    xs = torch.arange(W, device=x.device)[None, :]
    # the reference clamps the cut window's indices into the image, so the window is clipped at the borders
    y0, y1 = (oy - ch // 2).clamp(0, H - 1), (oy - ch // 2 + ch - 1).clamp(0, H - 1)
    x0, x1 = (ox - cw // 2).clamp(0, W - 1), (ox - cw // 2 + cw - 1).clamp(0, W - 1)
    inside = ((ys >= y0) & (ys <= y1))[:, :, None] & ((xs >= x0) & (xs <= x1))[:, None, :]
    return x * (~inside).to(x.dtype).unsqueeze(1)


AUGMENT_FNS = {
    'color': [rand_brightness, rand_saturation, rand_contrast],
    'translation': [rand_translation],
    'cutout': [rand_cutout],
}
