"""Fused bias + activation on libsg2b200, behind the reference op's Python API.

Public names follow thirdparty/stylegan3_ops/ops/bias_act.py: ``activation_funcs`` (:16-26, the table
``StyleGAN3/model.py:401`` reads ``def_gain`` from) and ``bias_act(x, b, dim, act, alpha, gain, clamp, impl)`` (:47).
Gradients of first and second order come from the same kernel run with ``grad = 1`` / ``grad = 2``
(reference :137-203).  Underneath: one prebuilt C-ABI kernel (``sg2_bias_act``), two autograd Functions
parameterised by a frozen config, no ``impl='ref'`` product path, no JIT plugin build.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from .. import _lib
from .._lib import amp_bwd, amp_fwd


class _Spec(dict):
    __getattr__ = dict.__getitem__


def _spec(idx, ref, second, alpha=0.0, gain=1.0):
    return _Spec(def_alpha=alpha, def_gain=gain, cuda_idx=idx, ref=ref, has_2nd_grad=second)


_SQRT2 = math.sqrt(2)
activation_funcs = {
    'linear':   _spec(1, '', False),
    'relu':     _spec(2, 'y', False, gain=_SQRT2),
    'lrelu':    _spec(3, 'y', False, alpha=0.2, gain=_SQRT2),
    'tanh':     _spec(4, 'y', True),
    'sigmoid':  _spec(5, 'y', True),
    'elu':      _spec(6, 'y', True),
    'selu':     _spec(7, 'y', True),
    'softplus': _spec(8, 'y', True),
    'swish':    _spec(9, 'x', True, gain=_SQRT2),
}


@dataclass(frozen=True)
class _Cfg:
    dim: int
    act: str
    alpha: float
    gain: float
    clamp: float        # < 0: disabled

    @property
    def spec(self):
        return activation_funcs[self.act]

    @property
    def trivial(self):
        return self.act == 'linear' and self.gain == 1 and self.clamp < 0


def _make_cfg(dim, act, alpha, gain, clamp) -> _Cfg:
    assert clamp is None or clamp >= 0
    s = activation_funcs[act]
    return _Cfg(int(dim), act, float(s.def_alpha if alpha is None else alpha),
                float(s.def_gain if gain is None else gain), float(-1 if clamp is None else clamp))


def _dense(x: torch.Tensor) -> torch.Tensor:
    """x as a dense buffer: channels_last stays channels_last (bias_act.py:140), else contiguous."""
    if x.ndim == 4 and x.stride(1) == 1 and x.is_contiguous(memory_format=torch.channels_last):
        return x
    return x.contiguous()


def _like(t: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """t laid out with ref's strides -- the kernel walks flat buffers (bias_act.cpp:10-22)."""
    if ref is None:
        return _dense(t)
    if t.stride() == ref.stride():
        return t
    out = torch.empty_strided(ref.shape, ref.stride(), dtype=t.dtype, device=t.device)
    return out.copy_(t)


def _kernel(cfg: _Cfg, x, b, xref, yref, dy, grad):
    lib = _lib.load()
    _lib.require_cuda(x)
    if x.dtype not in _lib.DTYPE_CODE:
        raise RuntimeError(f'bias_act: unsupported dtype {x.dtype}')
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    size_b, step_b = 0, 1
    if b is not None:
        if b.dtype != x.dtype or b.device != x.device:
            raise RuntimeError('bias_act: b must have the same dtype and device as x')
        size_b, step_b = b.numel(), x.stride(cfg.dim)
    _lib.check(lib.sg2_bias_act(
        x.data_ptr(), _lib.ptr(b), _lib.ptr(xref), _lib.ptr(yref), _lib.ptr(dy), y.data_ptr(),
        _lib.DTYPE_CODE[x.dtype], x.numel(), size_b, step_b, grad, cfg.spec.cuda_idx,
        cfg.alpha, cfg.gain, cfg.clamp, _lib.stream_ptr(x)), 'sg2_bias_act')
    return y


def _other_dims(t, dim):
    return [i for i in range(t.ndim) if i != dim]


class _BiasActFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, b, cfg: _Cfg):
        x = _dense(x)
        b = None if b is None else b.contiguous()
        y = x if (cfg.trivial and b is None) else _kernel(cfg, x, b, None, None, None, 0)
        need_x = 'x' in cfg.spec.ref or cfg.spec.has_2nd_grad
        ctx.cfg, ctx.has_b = cfg, b is not None
        # 'linear' has ref='' in the reference table, so its CUDA plugin never masks the gradient of a clamped
        # linear output; the reference's own 'ref' implementation (autograd through clamp) does.  Follow the
        # latter -- it is the mathematically correct one and what the oracle is pinned to.
        need_y = 'y' in cfg.spec.ref or (cfg.clamp >= 0 and 'x' not in cfg.spec.ref)
        ctx.save_for_backward(x if need_x else None, b if need_x else None, y if need_y else None)
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, dy):
        x, b, y = ctx.saved_tensors
        cfg = ctx.cfg
        dx = db = None
        if ctx.needs_input_grad[0] or (ctx.has_b and ctx.needs_input_grad[1]):
            dx = dy if cfg.trivial else _BiasActGradFn.apply(dy, x, b, y, cfg)
        if ctx.has_b and ctx.needs_input_grad[1]:
            db = dx.sum(_other_dims(dx, cfg.dim))
        return dx, db, None


class _BiasActGradFn(torch.autograd.Function):
    """dx = dy * act'(.) * gain (masked by the clamp), from the saved output y (or x for swish)."""

    @staticmethod
    @amp_fwd
    def forward(ctx, dy, x, b, y, cfg: _Cfg):
        ref = y if y is not None else x
        dy = _like(dy, ref)
        dx = _kernel(cfg, dy, b, x, y, None, 1)
        ctx.cfg = cfg
        ctx.save_for_backward(dy if cfg.spec.has_2nd_grad else None, x, b, y)
        return dx

    @staticmethod
    @amp_bwd
    def backward(ctx, d_dx):
        dy, x, b, y = ctx.saved_tensors
        cfg = ctx.cfg
        d_dy = d_x = d_b = None
        if ctx.needs_input_grad[0]:
            d_dy = _BiasActGradFn.apply(d_dx, x, b, y, cfg)
        if cfg.spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            ref = y if y is not None else x
            d_x = _kernel(cfg, _like(d_dx, ref), b, x, y, dy, 2)
            if ctx.needs_input_grad[2] and b is not None:
                d_b = d_x.sum(_other_dims(d_x, cfg.dim))
        return d_dy, d_x, d_b, None, None


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """``clamp(act(x + b) * gain)`` in one pass; arguments as thirdparty/stylegan3_ops/ops/bias_act.py:47-84."""
    assert isinstance(x, torch.Tensor)
    assert impl in ('ref', 'cuda')
    if impl == 'ref':
        raise RuntimeError("bias_act: impl='ref' is not a product path here; see oracle/ops_numpy.py")
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim
        assert b.shape[0] == x.shape[dim]
    return _BiasActFn.apply(x, b, _make_cfg(dim, act, alpha, gain, clamp))


def act_grad(gy: torch.Tensor, y: torch.Tensor, slope: float) -> torch.Tensor:
    """gy * (y > 0 ? 1 : slope): leaky-ReLU gradient from the saved OUTPUT; differentiable again in gy."""
    return _BiasActGradFn.apply(gy, None, None, y, _make_cfg(1, 'lrelu', slope, 1.0, None))
