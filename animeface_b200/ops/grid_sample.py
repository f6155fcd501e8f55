"""Image-space linear operators of the ADA pipeline on libsg2b200 (csrc/ada.cu), each closed under differentiation.

  affine_grid_sample(x, theta, size)   F.affine_grid(theta, size, align_corners=False) followed by
                                       grid_sample_gradfix.grid_sample(x, grid) -- bilinear, zeros padding
                                       (thirdparty/ada/augment.py:293-295, thirdparty/stylegan3_ops/ops/grid_sample_gradfix.py:19-77);
                                       the grid is evaluated inside the kernel, never stored
  reflect_pad(x, [x0, x1, y0, y1])     torch.nn.functional.pad(mode='reflect')                        (augment.py:284)
  color_affine(x, C)                   per-sample homogeneous colour transform C[:, :3, :3] x + C[:, :3, 3]  (augment.py:352-361)
Gradients flow to the image only (theta / C come from random draws); forward and adjoint are each other's backward, so
gradients of any order exist -- the property the reference's grid_sample_gradfix provides for R1.
"""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import amp_bwd, amp_fwd


def _f32(x):
    if x.dtype != torch.float32:
        raise RuntimeError('ADA ops: float32 only')
    _lib.require_cuda(x)
    return x.contiguous()


def _sample_raw(x, theta, in_hw, out_hw, adjoint):
    lib = _lib.load()
    x = _f32(x)
    n, c = x.shape[:2]
    (ih, iw), (oh, ow) = in_hw, out_hw
    y = torch.empty((n, c, ih, iw) if adjoint else (n, c, oh, ow), dtype=torch.float32, device=x.device)
    _lib.check(lib.sg2_affine_sample(x.data_ptr(), y.data_ptr(), theta.data_ptr(), n, c, ih, iw, oh, ow, 1 if adjoint else 0,
                                     _lib.stream_ptr(x)), 'sg2_affine_sample')
    return y


class _SampleFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, theta, out_hw, adjoint, in_hw):
        ctx.theta, ctx.out_hw, ctx.in_hw, ctx.adjoint = theta, out_hw, in_hw, adjoint
        return _sample_raw(x, theta, in_hw, out_hw, adjoint)

    @staticmethod
    @amp_bwd
    def backward(ctx, g):
        return _SampleFn.apply(g, ctx.theta, ctx.out_hw, not ctx.adjoint, ctx.in_hw), None, None, None, None


def affine_grid_sample(x, theta, size):
    """x [N,C,H,W], theta [N,2,3] (normalised output -> normalised input coordinates), size = output [N,C,H',W']."""
    assert x.ndim == 4 and tuple(theta.shape) == (x.shape[0], 2, 3) and len(size) == 4 and size[0] == x.shape[0] and size[1] == x.shape[1]
    theta = theta.detach().to(torch.float32).contiguous()
    return _SampleFn.apply(x, theta, (int(size[2]), int(size[3])), False, (int(x.shape[2]), int(x.shape[3])))


class _ReflectPadFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, pads, adjoint, hw):
        lib = _lib.load()
        x = _f32(x)
        n, c = x.shape[:2]
        h, w = hw
        px0, px1, py0, py1 = pads
        shape = (n, c, h, w) if adjoint else (n, c, h + py0 + py1, w + px0 + px1)
        y = torch.empty(shape, dtype=torch.float32, device=x.device)
        _lib.check(lib.sg2_reflect_pad(x.data_ptr(), y.data_ptr(), n * c, h, w, px0, px1, py0, py1, 1 if adjoint else 0,
                                       _lib.stream_ptr(x)), 'sg2_reflect_pad')
        ctx.pads, ctx.adjoint, ctx.hw = pads, adjoint, hw
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, g):
        return _ReflectPadFn.apply(g, ctx.pads, not ctx.adjoint, ctx.hw), None, None, None


def reflect_pad(x, pad):
    """pad = [x0, x1, y0, y1] (torch.nn.functional.pad order), each non-negative and smaller than the image."""
    pads = tuple(int(p) for p in pad)
    if not any(pads):
        return x
    return _ReflectPadFn.apply(x, pads, False, (int(x.shape[2]), int(x.shape[3])))


class _ColorFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, cmat, transpose):
        lib = _lib.load()
        x = _f32(x)
        n, c, h, w = x.shape
        assert c == 3
        y = torch.empty_like(x)
        _lib.check(lib.sg2_color_affine(x.data_ptr(), y.data_ptr(), cmat.data_ptr(), n, h * w, 1 if transpose else 0,
                                        _lib.stream_ptr(x)), 'sg2_color_affine')
        ctx.cmat, ctx.transpose = cmat, transpose
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, g):
        if ctx.transpose:
            # the adjoint op is y = M^T x; its derivative is M (the forward op without the offset)
            lin = ctx.cmat.clone()
            lin[:, :3, 3] = 0
            return _ColorFn.apply(g, lin, False), None, None
        return _ColorFn.apply(g, ctx.cmat, True), None, None


def color_affine(x, cmat):
    """x [N,3,H,W]; cmat [N,4,4]: y = cmat[:, :3, :3] @ x + cmat[:, :3, 3:]."""
    return _ColorFn.apply(x, cmat.detach().to(torch.float32).contiguous(), False)
