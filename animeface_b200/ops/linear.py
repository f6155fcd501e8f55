"""Fully connected layers of the reference path on libsg2b200 (csrc/linear.cu): no cuBLAS, no elementwise glue.

Replaces, for implementations/StyleGAN2/model.py:
  * ``ELR(nn.Linear)``                :29-37, 44-47   y = (x * coef) W^T + b
  * ``MapLinear`` + ``LeakyReLU``     :71-78, 272-279 y = lrelu(((x * coef) W^T + b) * lr)        (8 per Mapping.forward)
  * ``ModulatedConv2d.affine``        :102, 110       s = (w * coef) A^T + b (+ 1 added by the caller)
  * the discriminator epilogue        :392-396        Linear(8192, 512) -> LeakyReLU -> Linear(512, 1)
  * ``PixelNorm``                     :253-256
One launch per layer forward (scale, bias, gain and leaky-ReLU fused), two per layer backward (the leaky-ReLU gradient is
fused into their loads).  R1 differentiates twice through the discriminator epilogue (nnutils/loss/penalty.py:85-101): the
three kernels are a closed family under differentiation -- F(x,W) = x W^T, Dx(g,W) = g W, Dw(g,x) = g^T x -- exactly like the
convolution family of ops/conv2d.py, so gradients of any order exist.
"""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import amp_bwd, amp_fwd


def _f32c(t):
    if t.dtype != torch.float32:
        raise RuntimeError('linear: float32 only')
    return t.contiguous()


def _fwd_raw(x, w, b, coef, gain, slope):
    lib = _lib.load()
    _lib.require_cuda(x, w)
    x, w = _f32c(x), _f32c(w)
    B, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise RuntimeError(f'linear: input has {K} features, weight expects {w.shape[1]}')
    y = torch.empty((B, N), dtype=torch.float32, device=x.device)
    bb = None if b is None else _f32c(b.detach()).reshape(-1)
    nbytes = int(lib.sg2_linear_fwd_workspace(B, K, N))
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device) if nbytes else None
    _lib.check(lib.sg2_linear_fwd(x.data_ptr(), w.data_ptr(), _lib.ptr(bb), y.data_ptr(), B, K, N, float(coef), float(gain),
                                  float(1.0 if slope is None else slope), _lib.ptr(ws), _lib.stream_ptr(x)), 'sg2_linear_fwd')
    return y


def _dx_raw(gy, y, w, coef, gain, slope):
    lib = _lib.load()
    gy, w = _f32c(gy), _f32c(w)
    B, N = gy.shape
    K = w.shape[1]
    gx = torch.empty((B, K), dtype=torch.float32, device=gy.device)
    yy = None if (y is None or slope is None) else _f32c(y)
    _lib.check(lib.sg2_linear_bwd_data(gy.data_ptr(), _lib.ptr(yy), w.data_ptr(), gx.data_ptr(), B, K, N, float(coef), float(gain),
                                       float(1.0 if slope is None else slope), _lib.stream_ptr(gy)), 'sg2_linear_bwd_data')
    return gx


def _dw_raw(gy, y, x, coef, gain, slope, want_gb):
    lib = _lib.load()
    gy, x = _f32c(gy), _f32c(x)
    B, N = gy.shape
    K = x.shape[1]
    gw = torch.empty((N, K), dtype=torch.float32, device=gy.device)
    gb = torch.empty((N,), dtype=torch.float32, device=gy.device) if want_gb else None
    yy = None if (y is None or slope is None) else _f32c(y)
    _lib.check(lib.sg2_linear_bwd_weight(gy.data_ptr(), _lib.ptr(yy), x.data_ptr(), gw.data_ptr(), _lib.ptr(gb), B, K, N, float(coef),
                                         float(gain), float(1.0 if slope is None else slope), _lib.stream_ptr(gy)), 'sg2_linear_bwd_weight')
    return gw, gb


# ---- the closed family (any-order autograd); coef is the ELR constant ------------------------------------------------
class LinFn(torch.autograd.Function):
    """y = coef * x W^T"""

    @staticmethod
    @amp_fwd
    def forward(ctx, x, w, coef):
        ctx.coef = coef
        ctx.save_for_backward(x, w)
        return _fwd_raw(x, w, None, coef, 1.0, None)

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = LinDxFn.apply(gy, w, ctx.coef) if ctx.needs_input_grad[0] else None
        gw = LinDwFn.apply(gy, x, ctx.coef) if ctx.needs_input_grad[1] else None
        return gx, gw, None


class LinDxFn(torch.autograd.Function):
    """gx = coef * g W"""

    @staticmethod
    @amp_fwd
    def forward(ctx, g, w, coef):
        ctx.coef = coef
        ctx.save_for_backward(g, w)
        return _dx_raw(g, None, w, coef, 1.0, None)

    @staticmethod
    @amp_bwd
    def backward(ctx, gout):
        g, w = ctx.saved_tensors
        gg = LinFn.apply(gout, w, ctx.coef) if ctx.needs_input_grad[0] else None
        gw = LinDwFn.apply(g, gout, ctx.coef) if ctx.needs_input_grad[1] else None
        return gg, gw, None


class LinDwFn(torch.autograd.Function):
    """gw = coef * g^T x"""

    @staticmethod
    @amp_fwd
    def forward(ctx, g, x, coef):
        ctx.coef = coef
        ctx.save_for_backward(g, x)
        return _dw_raw(g, None, x, coef, 1.0, None, False)[0]

    @staticmethod
    @amp_bwd
    def backward(ctx, gout):
        g, x = ctx.saved_tensors
        gg = LinFn.apply(x, gout, ctx.coef) if ctx.needs_input_grad[0] else None
        gx = LinDxFn.apply(g, gout, ctx.coef) if ctx.needs_input_grad[1] else None
        return gg, gx, None


class LinearBiasActFn(torch.autograd.Function):
    """y = lrelu_slope(gain * (coef * x W^T + b)); fused forward, fused first-order backward; under create_graph the
    backward is composed from the closed family and the twice-differentiable activation gradient."""

    @staticmethod
    @amp_fwd
    def forward(ctx, x, w, b, coef, gain, slope):
        y = _fwd_raw(x, w, b, coef, gain, slope)
        ctx.coef, ctx.gain, ctx.slope = coef, gain, slope
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        if torch.is_grad_enabled():
            from .bias_act import act_grad
            gu = (act_grad(gy, y, ctx.slope) if ctx.slope is not None else gy) * ctx.gain
            gx = LinDxFn.apply(gu, w, ctx.coef) if need_x else None
            gw = LinDwFn.apply(gu, x, ctx.coef) if need_w else None
            gb = gu.sum(0) if need_b else None
            return gx, gw, gb, None, None, None
        gx = _dx_raw(gy, y, w, ctx.coef, ctx.gain, ctx.slope) if need_x else None
        gw = gb = None
        if need_w or need_b:
            gw, gb = _dw_raw(gy, y, x, ctx.coef, ctx.gain, ctx.slope, need_b)
        return gx, (gw if need_w else None), gb, None, None, None


def linear_bias_act(x, w, b=None, coef: float = 1.0, gain: float = 1.0, slope: float | None = None):
    """act(gain * ((x * coef) W^T + b)); x [B,K], w [N,K], b [N]; slope None = no activation."""
    lead = x.shape[:-1]
    y = LinearBiasActFn.apply(x.reshape(-1, x.shape[-1]), w, b, float(coef), float(gain), slope)
    return y.reshape(*lead, w.shape[0])


class PixelNormFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, eps):
        lib = _lib.load()
        _lib.require_cuda(x)
        x = _f32c(x)
        y = torch.empty_like(x)
        _lib.check(lib.sg2_pixelnorm(x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], float(eps), _lib.stream_ptr(x)), 'sg2_pixelnorm')
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        raise RuntimeError('pixel_norm: the fused kernel is for latents that do not require grad; use the composed expression')


def pixel_norm(x, eps: float = 1e-4):
    """x / (sqrt(mean_k x^2) + eps) (PixelNorm, model.py:253-256).  Latents never need a gradient on the training path; a
    tensor that does (or a non-2-D one) takes the differentiable expression."""
    if x.requires_grad or x.ndim != 2:
        return x / (x.pow(2).mean(dim=1, keepdim=True).sqrt() + eps)
    return PixelNormFn.apply(x, eps)
