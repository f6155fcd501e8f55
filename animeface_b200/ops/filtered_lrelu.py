"""``filtered_lrelu`` behind the reference's name and signature (thirdparty/stylegan3_ops/ops/filtered_lrelu.py:50-268).

bias -> zero-insert up-sampling by `up` + padding + FIR `fu` (gain up^2) -> leaky ReLU * gain, clamp -> FIR `fd` +
decimation by `down`: the alias-suppressed non-linearity of the StyleGAN3 generator (implementations/StyleGAN3/model.py:186-190).

Two executions of the same arithmetic (the reference's ``_filtered_lrelu_ref``, :121-147):

* **fused** (``csrc/filtered_lrelu.cu``, one launch): fp32 CUDA tensors (repacked to dense NCHW when they are not), a 1-D
  (separable) ``fu`` and a 1-D or 2-D ``fd``.
  The up-sampled intermediates live in shared memory; when a backward will follow, one byte per up-sampled element records
  the sign / clamp state.  The backward is the SAME kernel reading that mask instead of applying the activation, with the
  filters in swapped roles -- the construction of the reference's ``FilteredLReluCuda.backward`` (:221-252), re-derived on the
  forward's own up-sampled grid; a 2-D ``fd`` (the radial filters) becomes the 2-D first stage of that launch.
* **composed** (``bias_act`` / ``upfirdn2d`` launches, the up-sampled tensor materialised): everything else -- other dtypes,
  a 2-D ``fu``, and any backward under ``create_graph`` (gradients of any order exist through the ops' own
  closed families).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._lib import amp_bwd, amp_fwd
from . import bias_act as _ba
from . import upfirdn2d as _up
from .upfirdn2d import _quad, _taps

fused_enabled = True          # module switch for A/B tests


def _composed(x, fu, fd, b, up, down, pads, gain, slope, clamp, flip_filter):
    px0, px1, py0, py1 = pads
    y = _ba.bias_act(x=x, b=b) if b is not None else x
    y = _up.upfirdn2d(x=y, f=fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    y = _ba.bias_act(x=y, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    return _up.upfirdn2d(x=y, f=fd, down=down, flip_filter=flip_filter)


def _corr(f, flip_filter, device):
    """The filter as the correlation kernel upfirdn2d applies (it convolves unless flip_filter): float32, contiguous."""
    if f is None:
        return torch.ones(1, dtype=torch.float32, device=device)
    f = f.detach().to(device=device, dtype=torch.float32)
    if not flip_filter:
        f = f.flip(list(range(f.ndim)))
    return f.contiguous()


def _launch(x, b, y, mask, fu_c, fd_c, up, pad0, z_hw, down, doff, up_gain, gain, slope, clamp, mode):
    lib = _lib.load()
    n, c, in_h, in_w = x.shape
    out_h, out_w = y.shape[2:]
    _lib.check(lib.sg2_filtered_lrelu(
        x.data_ptr(), _lib.ptr(b), y.data_ptr(), _lib.ptr(mask), fu_c.data_ptr(), fd_c.data_ptr(), 1 if fu_c.ndim == 2 else 0,
        1 if fd_c.ndim == 2 else 0,
        n * c, c, in_h, in_w, up, pad0[0], pad0[1], z_hw[0], z_hw[1], fu_c.shape[-1], fd_c.shape[-1], down, doff[0], doff[1],
        out_h, out_w, float(up_gain), float(gain), float(slope), float(-1.0 if clamp is None else clamp), mode,
        _lib.stream_ptr(x)), 'sg2_filtered_lrelu')


def _fusable(x, fu, fd, b):
    if not (fused_enabled and x.is_cuda and x.dtype == torch.float32 and x.numel() > 0):
        return False
    if fu is not None and (fu.ndim != 1 or fu.shape[0] > 64):
        return False
    if fd is not None and (fd.shape[-1] > 32 or (fd.ndim == 2 and fd.shape[0] != fd.shape[1])):
        return False
    return x.shape[0] * x.shape[1] <= 65535


class FilteredLReluFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, b, fu, fd, up, down, pads, gain, slope, clamp, flip_filter):
        px0, px1, py0, py1 = pads
        n, c, in_h, in_w = x.shape
        fu_c, fd_c = _corr(fu, flip_filter, x.device), _corr(fd, flip_filter, x.device)
        fu_n, fd_n = fu_c.shape[-1], fd_c.shape[-1]
        zw, zh = in_w * up + px0 + px1 - (fu_n - 1), in_h * up + py0 + py1 - (fu_n - 1)
        out_w, out_h = (zw - (fd_n - 1) + (down - 1)) // down, (zh - (fd_n - 1) + (down - 1)) // down
        if min(zw, zh, out_w, out_h) <= 0:
            raise RuntimeError('filtered_lrelu: the padding leaves no output')
        y = torch.empty((n, c, out_h, out_w), dtype=torch.float32, device=x.device)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        mask = torch.empty((n, c, zh, zw), dtype=torch.uint8, device=x.device) if need_grad else None
        bb = None if b is None else b.detach().to(torch.float32).contiguous()
        _launch(x, bb, y, mask, fu_c, fd_c, up, (px0, py0), (zh, zw), down, (0, 0), up ** 2, gain, slope, clamp, 1 if need_grad else 0)
        ctx.cfg = (up, down, pads, gain, slope, clamp, flip_filter, (zh, zw), (in_h, in_w))
        ctx.has_b = b is not None
        # x and b are kept only for the create_graph path (the first-order backward needs the mask and the filters)
        ctx.save_for_backward(x, b if b is not None else torch.empty(0, device=x.device), fu_c, fd_c, mask if need_grad else torch.empty(0, device=x.device),
                              fu if fu is not None else torch.empty(0, device=x.device), fd if fd is not None else torch.empty(0, device=x.device))
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, dy):
        x, b, fu_c, fd_c, mask, fu, fd = ctx.saved_tensors
        up, down, pads, gain, slope, clamp, flip_filter, (zh, zw), (in_h, in_w) = ctx.cfg
        b = b if ctx.has_b else None
        if torch.is_grad_enabled():
            # create_graph: differentiate the composed form of the same function (every piece stays differentiable)
            with torch.enable_grad():
                xd = x.detach().requires_grad_(True)
                bd = None if b is None else b.detach().requires_grad_(True)
                yc = _composed(xd, fu if fu.numel() else None, fd if fd.numel() else None, bd, up, down, pads, gain, slope, clamp, flip_filter)
                ins = [xd] + ([bd] if bd is not None else [])
                gs = torch.autograd.grad(yc, ins, dy, create_graph=True, allow_unused=True)
            return gs[0], (gs[1] if bd is not None else None), None, None, None, None, None, None, None, None, None
        px0, _, py0, _ = pads
        n, c = x.shape[:2]
        fu_n, fd_n = fu_c.shape[-1], fd_c.shape[-1]
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty((n, c, in_h, in_w), dtype=torch.float32, device=dy.device)
        g_fu = fu_c.flip(0).contiguous()                       # second stage of the backward: fu, flipped, decimating by `up`
        doff = (px0 - (fu_n - 1), py0 - (fu_n - 1))
        # one launch: dy -> (x `down`, fd flipped: separable or, for the radial filters, the full 2-D adjoint) -> * mask -> (fu flipped, / `up`)
        g_fd = fd_c.flip(list(range(fd_c.ndim))).contiguous()
        _launch(dy, None, dx, mask, g_fd, g_fu, down, (fd_n - 1, fd_n - 1), (zh, zw), up, doff, 1.0, gain * up ** 2, slope, None, 2)
        db = dx.sum((0, 2, 3)).to(b.dtype) if (b is not None and ctx.needs_input_grad[1]) else None
        return (dx if ctx.needs_input_grad[0] else None), db, None, None, None, None, None, None, None, None, None


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
    pads = _quad(padding)
    px0, px1, py0, py1 = pads
    fu_w, fu_h = _taps(fu)
    fd_w, fd_h = _taps(fd)
    n, c, in_h, in_w = x.shape
    out_w = (in_w * up + (px0 + px1) - (fu_w - 1) - (fd_w - 1) + (down - 1)) // down
    out_h = (in_h * up + (py0 + py1) - (fu_h - 1) - (fd_h - 1) + (down - 1)) // down
    if _fusable(x, fu, fd, b):
        # the kernel walks (sample, channel) planes: a channels_last or cropped input is repacked first (one pass over the SMALL tensor;
        # the composed form would write and re-read the up^2-times larger intermediates instead)
        y = FilteredLReluFn.apply(x.contiguous(), b, fu, fd, up, down, pads, float(gain), float(slope), None if clamp is None else float(clamp), bool(flip_filter))
    else:
        y = _composed(x, fu, fd, b, up, down, pads, gain, slope, clamp, flip_filter)
    assert tuple(y.shape) == (n, c, out_h, out_w), (tuple(y.shape), (n, c, out_h, out_w))
    return y
