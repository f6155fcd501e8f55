"""StyleGAN2 resampling on libsg2b200: fused bilinear-x2 (+blur) and 2x2 average pooling.

Replaces, for the reference model:
  * ``Upsample2x('bilinear')`` + ``Blur2d``     implementations/StyleGAN2/model.py:56-58, 138-149, 160-161
  * bare ``Upsample2x('bilinear')`` of ToImage   implementations/StyleGAN2/model.py:243, 248-249
  * ``Downsample2x('avg')`` + ``(x + t)/sqrt(2)``  implementations/StyleGAN2/model.py:61-63, 209-212
Both operators are linear, so each Function's backward is the other member of its (op, adjoint)
pair and gradients of any order exist.
"""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import amp_bwd, amp_fwd
from .conv2d import _cl, _empty_cl


def _is_cl(x):
    return x.shape[1] % 4 == 0


def _up2x(x, scale, blur, adjoint):
    """fwd: [n,c,h,w] -> [n,c,2h,2w]; adjoint: [n,c,2h,2w] -> [n,c,h,w]."""
    lib = _lib.load()
    _lib.require_cuda(x)
    if x.dtype != torch.float32:
        raise RuntimeError('up2x: float32 only')
    n, c, h, w = x.shape
    nhwc = _is_cl(x)
    if adjoint:
        assert h % 2 == 0 and w % 2 == 0
        h, w = h // 2, w // 2
        oh, ow = h, w
    else:
        oh, ow = 2 * h, 2 * w
    if nhwc:
        x = _cl(x)
        y = _empty_cl(n, c, oh, ow, x)
    else:
        x = x.contiguous()
        y = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
    sc = None if scale is None else scale.detach().to(torch.float32).contiguous()
    fn = lib.sg2_up2x_adj if adjoint else lib.sg2_up2x_fwd
    _lib.check(fn(x.data_ptr(), y.data_ptr(), _lib.ptr(sc), n, c, h, w, int(nhwc), int(blur), _lib.stream_ptr(x)),
               'sg2_up2x_adj' if adjoint else 'sg2_up2x_fwd')
    return y


class Up2xFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, blur):
        ctx.blur = blur
        return _up2x(x, None, blur, False)

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        return Up2xAdjFn.apply(gy, ctx.blur), None


class Up2xAdjFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, gy, blur):
        ctx.blur = blur
        return _up2x(gy, None, blur, True)

    @staticmethod
    @amp_bwd
    def backward(ctx, g):
        return Up2xFn.apply(g, ctx.blur), None


def upsample2x_blur(x):
    """nn.Upsample(x2, bilinear, align_corners=False) followed by Blur2d -- one kernel."""
    return Up2xFn.apply(x, True)


def upsample2x_bilinear(x):
    """nn.Upsample(x2, bilinear, align_corners=False)."""
    return Up2xFn.apply(x, False)


def _avgpool(x, t, alpha, adjoint, t_pooled=False, want_signs=False):
    lib = _lib.load()
    _lib.require_cuda(x)
    n, c, h, w = x.shape
    x = _cl(x)
    if adjoint:
        y = _empty_cl(n, c, 2 * h, 2 * w, x)
        _lib.check(lib.sg2_avgpool2_adj(x.data_ptr(), y.data_ptr(), float(alpha), n, c, 2 * h, 2 * w,
                                        _lib.stream_ptr(x)), 'sg2_avgpool2_adj')
    else:
        t = None if t is None else _cl(t)
        y = _empty_cl(n, c, h // 2, w // 2, x)
        if t is not None and tuple(t.shape) != ((n, c, h // 2, w // 2) if t_pooled else (n, c, h, w)):
            raise RuntimeError(f'avgpool2: residual input has shape {tuple(t.shape)}')
        signs = torch.empty((n, h // 2, w // 2, c // 4), dtype=torch.int16, device=x.device) if want_signs else None
        _lib.check(lib.sg2_avgpool2_fwd(x.data_ptr(), _lib.ptr(t), y.data_ptr(), _lib.ptr(signs), float(alpha), n, c, h, w,
                                        1 if t_pooled else 0, _lib.stream_ptr(x)), 'sg2_avgpool2_fwd')
        if want_signs:
            return y, signs
    return y


class AvgPool2Fn(torch.autograd.Function):
    """y = alpha * (avg2x2(x) + avg2x2(t)); t may be None."""

    @staticmethod
    @amp_fwd
    def forward(ctx, x, t, alpha):
        ctx.alpha, ctx.has_t = alpha, t is not None
        return _avgpool(x, t, alpha, False)

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        g = AvgPool2AdjFn.apply(gy, ctx.alpha)
        return g, (g if ctx.has_t else None), None


class AvgPool2AdjFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, gy, alpha):
        ctx.alpha = alpha
        return _avgpool(gy, None, alpha, True)

    @staticmethod
    @amp_bwd
    def backward(ctx, g):
        return AvgPool2Fn.apply(g, None, ctx.alpha), None


def avgpool2(x, t=None, alpha=1.0):
    """AvgPool2d(2) of x (and t), summed and scaled: DBlock's ``(down(x) + down(t)) / sqrt(2)``."""
    if x.shape[1] % 4 != 0:
        raise RuntimeError('avgpool2: channel count must be a multiple of 4')
    return AvgPool2Fn.apply(x, t, float(alpha))
