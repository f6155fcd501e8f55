"""upfirdn2d on libsg2b200, behind the reference op's Python API.

Public names and argument meaning follow thirdparty/stylegan3_ops/ops/upfirdn2d.py
(``setup_filter`` :64, ``upfirdn2d`` :112, ``filter2d`` :271, ``upsample2d`` :307, ``downsample2d`` :346) so
that code written against the reference op runs unchanged.  What is different underneath:
  * one prebuilt C-ABI kernel (``sg2_upfirdn2d``) instead of a JIT-built pybind plugin (custom_ops.py:53);
  * a single autograd.Function parameterised by a frozen ``_Plan`` instead of a class per argument tuple;
    the backward of a plan is the plan with up/down swapped, mirrored padding and the filter flipped
    (upfirdn2d.py:245-263), so gradients of any order exist;
  * no ``impl='ref'`` product path -- the CPU restatement is test infrastructure (oracle/ops_numpy.py).
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import torch

from .. import _lib
from .._lib import amp_bwd, amp_fwd


def _pair(v, what):
    """int or [x, y] -> (x, y) of positive ints."""
    if isinstance(v, int):
        v = (v, v)
    assert isinstance(v, (list, tuple)) and len(v) == 2 and all(isinstance(e, int) for e in v), what
    assert v[0] >= 1 and v[1] >= 1, what
    return int(v[0]), int(v[1])


def _quad(v):
    """int, [x, y] or [x0, x1, y0, y1] -> (x0, x1, y0, y1)."""
    if isinstance(v, int):
        return v, v, v, v
    assert isinstance(v, (list, tuple)) and all(isinstance(e, int) for e in v)
    if len(v) == 2:
        return v[0], v[0], v[1], v[1]
    assert len(v) == 4
    return tuple(v)


def _taps(f):
    """(fw, fh) of a filter tensor (None = identity)."""
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Build the float32 FIR tensor ``upfirdn2d`` expects (reference: upfirdn2d.py:64-108).

    1-D taps become an outer-product 2-D filter unless ``separable`` (default: 1-D and >= 8 taps)."""
    t = torch.as_tensor(1 if f is None else f, dtype=torch.float32)
    assert t.ndim <= 2 and t.numel() > 0
    t = t.reshape(1) if t.ndim == 0 else t
    sep = (t.ndim == 1 and t.numel() >= 8) if separable is None else separable
    if t.ndim == 1 and not sep:
        t = torch.outer(t, t)
    assert t.ndim == (1 if sep else 2)
    if normalize:
        t = t / t.sum()
    if flip_filter:
        t = t.flip(tuple(range(t.ndim)))
    return (t * gain ** (t.ndim / 2)).to(device=device)


@dataclass(frozen=True)
class _Plan:
    upx: int = 1
    upy: int = 1
    downx: int = 1
    downy: int = 1
    px0: int = 0
    px1: int = 0
    py0: int = 0
    py1: int = 0
    flip: bool = False
    gain: float = 1.0

    def out_size(self, ih, iw, fh, fw):
        ow = (iw * self.upx + self.px0 + self.px1 - fw + self.downx) // self.downx
        oh = (ih * self.upy + self.py0 + self.py1 - fh + self.downy) // self.downy
        return oh, ow

    def adjoint(self, ih, iw, oh, ow, fh, fw):
        return _Plan(self.downx, self.downy, self.upx, self.upy,
                     fw - self.px0 - 1, iw * self.upx - ow * self.downx + self.px0 - self.upx + 1,
                     fh - self.py0 - 1, ih * self.upy - oh * self.downy + self.py0 - self.upy + 1,
                     not self.flip, self.gain)


def _run(x: torch.Tensor, f2d: torch.Tensor, plan: _Plan) -> torch.Tensor:
    """One launch of sg2_upfirdn2d; the output keeps x's memory format (upfirdn2d.cpp:32)."""
    lib = _lib.load()
    n, c, ih, iw = x.shape
    fh, fw = f2d.shape
    oh, ow = plan.out_size(ih, iw, fh, fw)
    if ow < 1 or oh < 1:
        raise RuntimeError('upfirdn2d: output must be at least 1x1')
    fmt = torch.channels_last if (x.stride(1) == 1 and c > 1) else torch.contiguous_format
    y = torch.empty((n, c, oh, ow), dtype=x.dtype, device=x.device, memory_format=fmt)
    f2d = f2d.contiguous()
    _lib.check(lib.sg2_upfirdn2d(
        x.data_ptr(), f2d.data_ptr(), y.data_ptr(), _lib.DTYPE_CODE[x.dtype], n, c, ih, iw, _lib.strides4(x),
        oh, ow, _lib.strides4(y), fh, fw, plan.upx, plan.upy, plan.downx, plan.downy,
        plan.px0, plan.px1, plan.py0, plan.py1, int(plan.flip), float(plan.gain), _lib.stream_ptr(x)), 'sg2_upfirdn2d')
    return y


class _Upfirdn2dFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, f, plan: _Plan):
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        _lib.require_cuda(x)
        if x.dtype not in _lib.DTYPE_CODE:
            raise RuntimeError(f'upfirdn2d: unsupported dtype {x.dtype}')
        if x.numel() == 0:
            raise RuntimeError('upfirdn2d: x has zero size')
        if f is None:
            f = torch.ones((1, 1), dtype=torch.float32, device=x.device)
        if f.ndim == 1 and f.shape[0] == 1:
            f = f.square().unsqueeze(0)          # separable 1-tap == full 1x1 (upfirdn2d.py:229-230)
        assert f.ndim in (1, 2)
        if f.dtype != torch.float32:
            raise RuntimeError('upfirdn2d: f must be float32')
        if f.device != x.device:
            raise RuntimeError('upfirdn2d: f must reside on the same device as x')
        if f.ndim == 2:
            y = _run(x, f, plan)
        else:
            # separable taps: a horizontal then a vertical 1-D pass, the gain rides on the second
            horiz = replace(plan, upy=1, downy=1, py0=0, py1=0, gain=1.0)
            vert = replace(plan, upx=1, downx=1, px0=0, px1=0)
            y = _run(_run(x, f.unsqueeze(0), horiz), f.unsqueeze(1), vert)
        ctx.save_for_backward(f)
        ctx.plan, ctx.in_hw = plan, (x.shape[2], x.shape[3])
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, dy):
        f, = ctx.saved_tensors
        assert not ctx.needs_input_grad[1], 'upfirdn2d: the filter is not differentiable'
        if not ctx.needs_input_grad[0]:
            return None, None, None
        fw, fh = _taps(f)
        adj = ctx.plan.adjoint(ctx.in_hw[0], ctx.in_hw[1], dy.shape[2], dy.shape[3], fh, fw)
        return _Upfirdn2dFn.apply(dy, f, adj), None, None


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Pad -> zero-insert upsample -> FIR -> decimate (reference: upfirdn2d.py:112-157)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ('ref', 'cuda')
    if impl == 'ref':
        raise RuntimeError("upfirdn2d: impl='ref' is not a product path here; see oracle/ops_numpy.py")
    upx, upy = _pair(up, 'up')
    downx, downy = _pair(down, 'down')
    px0, px1, py0, py1 = _quad(padding)
    return _Upfirdn2dFn.apply(x, f, _Plan(upx, upy, downx, downy, px0, px1, py0, py1, bool(flip_filter), float(gain)))


def _centered(padding, fw, fh, lo_x, hi_x, lo_y, hi_y):
    px0, px1, py0, py1 = _quad(padding)
    return [px0 + lo_x, px1 + hi_x, py0 + lo_y, py1 + hi_y]


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Same-size FIR filtering (reference: upfirdn2d.py:271-303)."""
    fw, fh = _taps(f)
    p = _centered(padding, fw, fh, fw // 2, (fw - 1) // 2, fh // 2, (fh - 1) // 2)
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Integer-factor upsampling, output = input * up (reference: upfirdn2d.py:307-342)."""
    ux, uy = _pair(up, 'up')
    fw, fh = _taps(f)
    p = _centered(padding, fw, fh, (fw + ux - 1) // 2, (fw - ux) // 2, (fh + uy - 1) // 2, (fh - uy) // 2)
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * ux * uy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Integer-factor downsampling, output = input / down (reference: upfirdn2d.py:346-381)."""
    dx, dy = _pair(down, 'down')
    fw, fh = _taps(f)
    p = _centered(padding, fw, fh, (fw - dx + 1) // 2, (fw - dx) // 2, (fh - dy + 1) // 2, (fh - dy) // 2)
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
