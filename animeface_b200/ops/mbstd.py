"""Minibatch standard deviation on libsg2b200.

Replaces ``MiniBatchStdDev.forward`` implementations/StyleGAN2/model.py:215-236 (same math as
nnutils/module/layers.py:40-52): 2 launches instead of ~8 ATen kernels + ``torch.cat``.
First-order backward is a kernel; the backward of that backward (R1 differentiates through this
layer: nnutils/loss/penalty.py:85-101) is expressed with differentiable torch ops on the tiny
[B,C,4,4] tensor -- it runs on the GPU, and only on the 1-in-16 R1 steps.
"""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import amp_bwd, amp_fwd


def _resolve_groups(batch, group_size):
    return group_size if batch % group_size == 0 else batch       # model.py:234-236


def _bwd_formula(x, gy, G, eps):
    """Differentiable restatement of sg2_mbstd_bwd (used only when a graph of the backward is needed)."""
    n, c, h, w = x.shape
    M = n // G
    xg = x.reshape(G, M, c, h, w)
    mu = xg.mean(0, keepdim=True)
    dlt = xg - mu
    sd = (dlt.square().mean(0, keepdim=True) + eps).sqrt()
    gf = gy[:, c].reshape(G, M, h * w).sum((0, 2))                              # [M]
    coeff = gf.reshape(1, M, 1, 1, 1) / (G * c * h * w)
    return gy[:, :c] + (coeff * dlt / sd).reshape(n, c, h, w)


class MbstdFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, G, eps):
        lib = _lib.load()
        _lib.require_cuda(x)
        if x.dtype != torch.float32:
            raise RuntimeError('mbstd: float32 only')
        n, c, h, w = x.shape
        cl = x.stride(1) == 1
        y = torch.empty((n, c + 1, h, w), dtype=torch.float32, device=x.device,
                        memory_format=torch.channels_last if cl else torch.contiguous_format)
        stat = torch.empty(n // G, dtype=torch.float32, device=x.device)
        _lib.check(lib.sg2_mbstd_fwd(x.data_ptr(), _lib.strides4(x), y.data_ptr(), _lib.strides4(y), stat.data_ptr(),
                                     n, c, h, w, G, float(eps), _lib.stream_ptr(x)), 'sg2_mbstd_fwd')
        ctx.G, ctx.eps = G, eps
        ctx.save_for_backward(x)
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        x, = ctx.saved_tensors
        return MbstdGradFn.apply(x, gy, ctx.G, ctx.eps), None, None


class MbstdGradFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, gy, G, eps):
        lib = _lib.load()
        n, c, h, w = x.shape
        gx = torch.empty_like(x)
        _lib.check(lib.sg2_mbstd_bwd(x.data_ptr(), _lib.strides4(x), gy.data_ptr(), _lib.strides4(gy), gx.data_ptr(),
                                     _lib.strides4(gx), n, c, h, w, G, float(eps), _lib.stream_ptr(x)), 'sg2_mbstd_bwd')
        ctx.G, ctx.eps = G, eps
        ctx.save_for_backward(x, gy)
        return gx

    @staticmethod
    @amp_bwd
    def backward(ctx, ggx):
        x, gy = ctx.saved_tensors
        outer = torch.is_grad_enabled()
        with torch.enable_grad():
            xd = x.detach().requires_grad_(True)
            gd = gy.detach().requires_grad_(True)
            out = _bwd_formula(xd, gd, ctx.G, ctx.eps)
            gx, ggy = torch.autograd.grad(out, (xd, gd), ggx, create_graph=outer)
        return gx, ggy, None, None


def minibatch_stddev(x, group_size, eps=1e-4):
    """[B,C,H,W] -> [B,C+1,H,W]; group handling as model.py:221-236."""
    return MbstdFn.apply(x, _resolve_groups(x.shape[0], group_size), float(eps))
