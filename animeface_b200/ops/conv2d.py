"""Dense and modulated 2-D convolution on libsg2b200 (stride 1, 'same' zero padding, k in {1, 3}).

Replaces, for the StyleGAN2 path of the reference:
  * ``nn.Conv2d`` inside ``ELR``                    implementations/StyleGAN2/model.py:29-37, 50-53
  * ``ModulatedConv2d.forward``                      implementations/StyleGAN2/model.py:106-132
  * their autograd (``convolution_backward``), composed as a closed family of three ops --
    forward conv, data-gradient conv, weight-gradient -- exactly the way
    thirdparty/stylegan3_ops/ops/conv2d_gradfix.py:99-187 composes them, so gradients of arbitrary
    order exist (R1 needs the second order: nnutils/loss/penalty.py:11-26, 85-101).

Layout: activations are logical NCHW tensors in ``torch.channels_last`` memory format (NHWC in HBM);
weights keep the reference layout ``[co, ci, k, k]`` and are packed per call by the library.
"""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import amp_bwd, amp_fwd

import os

IMPL_AUTO, IMPL_SIMT, IMPL_HALO, IMPL_HALO32 = 0, 1, 4, 5
_default_impl = int(os.environ.get('SG2_CONV_IMPL', '0'))     # 0 = auto; see set_default_impl
# experiment switch (profiles/r2_*): data-gradient / second-order convolutions on the fp32-class kernel instead of bf16x3
_grad_precise = os.environ.get('SG2_GRAD_PRECISE', '0') == '1'

# bench.py sets this to a list to time every convolution launch with CUDA events on the launching stream:
# entries are (kind, flops, start_event, end_event).  None = no timing (the default).
launch_log = None


class _timed:
    def __init__(self, kind, flops, label=''):
        self.kind, self.flops, self.label = kind, flops, label

    def __enter__(self):
        if launch_log is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if launch_log is not None:
            self.e1.record()
            launch_log.append((self.kind, self.flops, self.e0, self.e1, self.label))


def set_default_impl(impl: int) -> None:
    """0 = auto (tcgen05 where the shape allows -- fp32-class split operands for forward convs, bf16x3 for gradients; halo
    kernels where the image tiles by 8x16 -- else fp32 SIMT), 1 = fp32 kernels everywhere, 4 / 5 = prefer the bf16x3 /
    fp16x3 + promotion halo kernel wherever it applies."""
    global _default_impl
    assert impl in (0, 1, 4, 5)
    _default_impl = impl


def _cl(x: torch.Tensor) -> torch.Tensor:
    """Dense channels_last view of a 4-D tensor (copy only if needed)."""
    n, c, h, w = x.shape
    want = (h * w * c, 1, w * c, c)
    if x.stride() == want:
        return x
    # torch treats size-1 dims as "any stride": normalise explicitly.
    y = torch.empty_strided((n, c, h, w), want, dtype=x.dtype, device=x.device)
    y.copy_(x)
    return y


def _empty_cl(n, c, h, w, like: torch.Tensor) -> torch.Tensor:
    return torch.empty_strided((n, c, h, w), (h * w * c, 1, w * c, c), dtype=torch.float32, device=like.device)


def _pack(w: torch.Tensor, coef: float, transpose: bool, impl: int) -> torch.Tensor:
    lib = _lib.load()
    co, ci, k, _ = w.shape
    nbytes = lib.sg2_conv2d_packed_size(co, ci, k, impl)
    if nbytes < 0:
        raise RuntimeError(f'conv2d: unsupported weight shape {tuple(w.shape)}')
    buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    wc = w.detach().contiguous()
    _lib.check(lib.sg2_conv2d_pack_weight(wc.data_ptr(), buf.data_ptr(), co, ci, k, float(coef),
                                          1 if transpose else 0, impl, _lib.stream_ptr(w)), 'sg2_conv2d_pack_weight')
    return buf


def _conv_raw(x, w, coef, transpose, in_scale=None, out_scale=None, bias=None, noise=None,
              slope=None, out_nchw=False, impl=None, precise=None, gain=1.0):
    """One library call: y = act(out_scale * conv(x * in_scale, w*coef) + bias + noise).

    transpose=True runs the data-gradient conv (x has w.shape[0] channels, y has w.shape[1]).
    precise: fp32-class tensor-core kernel (big/small split + promotion) instead of bf16x3.  Default: yes for forward
    convolutions (their outputs decide leaky-ReLU signs), no for data gradients."""
    strict = impl is not None                      # an explicit request must be honoured or fail loudly
    impl = _default_impl if impl is None else impl
    _lib.require_cuda(x, w)
    lib = _lib.load()
    co, ci, k, _ = w.shape
    cin, cout = (co, ci) if transpose else (ci, co)
    n, cx, h, wd = x.shape
    precise = (not transpose or _grad_precise) if precise is None else (precise or _grad_precise)
    req = impl
    impl = lib.sg2_conv2d_select_impl(n, h, wd, cin, cout, k, req, 1 if precise else 0)
    if impl < 0 and not strict:
        impl = IMPL_SIMT                            # a global tensor-core preference falls back to the fp32 SIMT kernel
    if impl < 0:
        raise RuntimeError(f'conv2d: implementation {req} does not take n={n} h={h} w={wd} ci={cin} co={cout} k={k}')
    if cx != cin:
        raise RuntimeError(f'conv2d: input has {cx} channels, weight expects {cin}')
    if x.dtype != torch.float32 or w.dtype != torch.float32:
        raise RuntimeError('conv2d: float32 only')
    x = _cl(x)
    packed = _pack(w, coef, transpose, impl)
    if out_nchw:
        y = torch.empty((n, cout, h, wd), dtype=torch.float32, device=x.device)
    else:
        y = _empty_cl(n, cout, h, wd, x)
    f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
    in_scale, out_scale, bias, noise = f32(in_scale), f32(out_scale), f32(bias), f32(noise)
    with _timed('dgrad' if transpose else 'fwd', 2.0 * n * h * wd * cin * cout * k * k, f'{cin}->{cout} k{k} @{h}x{wd} n{n} impl{impl}'):
        _lib.check(lib.sg2_conv2d_fwd(
            x.data_ptr(), packed.data_ptr(), y.data_ptr(), _lib.strides4(y), n, h, wd, cin, cout, k,
            _lib.ptr(in_scale), _lib.ptr(out_scale), _lib.ptr(bias), _lib.ptr(noise),
            3 if slope is not None else 1, float(slope if slope is not None else 0.0), float(gain),
            impl, _lib.stream_ptr(x)), 'sg2_conv2d_fwd')
    return y


def _wgrad_raw(x, gy, k, coef, in_scale=None, out_scale=None, impl=None):
    if impl is None:
        impl = IMPL_SIMT if _default_impl == IMPL_SIMT else IMPL_AUTO
    _lib.require_cuda(x, gy)
    lib = _lib.load()
    n, ci, h, wd = x.shape
    co = gy.shape[1]
    x, gy = _cl(x), _cl(gy)
    dw = torch.empty((co, ci, k, k), dtype=torch.float32, device=x.device)
    f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
    in_scale, out_scale = f32(in_scale), f32(out_scale)
    ws = _workspace(lib.sg2_conv2d_wgrad_workspace(n, h, wd, ci, co, k, impl), x.device, 'sg2_conv2d_wgrad_workspace')
    with _timed('wgrad', 2.0 * n * h * wd * ci * co * k * k, f'{ci}->{co} k{k} @{h}x{wd} n{n} fp32-operands'):
        _lib.check(lib.sg2_conv2d_wgrad(x.data_ptr(), gy.data_ptr(), dw.data_ptr(), n, h, wd, ci, co, k, float(coef),
                                        _lib.ptr(in_scale), _lib.ptr(out_scale), 0, impl, ws.data_ptr(), _lib.stream_ptr(x)),
                   'sg2_conv2d_wgrad')
    return dw


# ----------------------------------------------------------------------------------------------
# bf16 pair planes (csrc/planes.cu): the first-order backward fast path.  A gradient tensor / activation is written ONCE
# as hi = bf16(v), lo = bf16(v - hi) ([2][n,h,w,c], 4 bytes per element) by the elementwise pass that touches it anyway;
# the tcgen05 data-gradient and weight-gradient kernels then take their operands by TMA straight into swizzled tiles.

planes_enabled = True           # tests / A-B measurements flip this to compare with the fp32-operand kernels


def _workspace(nbytes: int, device, what: str) -> torch.Tensor:
    """Scratch for the deterministic two-pass reductions (every split stores its partial, a second kernel adds them in a
    fixed order).  Allocated per call from torch's caching allocator, so it is stream-ordered and capture-safe."""
    if nbytes < 0:
        raise RuntimeError(f'{what}: unsupported shape')
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _planes_ok(n, h, wd, cin, cout, k, wgrad: bool) -> bool:
    return planes_enabled and bool(_lib.load().sg2_conv2d_planes_supported(n, h, wd, cin, cout, k, 1 if wgrad else 0))


def _split_planes(x, scale=None):
    """planes[2][n,h,w,c] = split(x * scale[b,c]); x channels_last fp32."""
    lib = _lib.load()
    n, c, h, wd = x.shape
    x = _cl(x)
    planes = torch.empty((2, n, h, wd, c), dtype=torch.bfloat16, device=x.device)
    sc = None if scale is None else scale.detach().to(torch.float32).contiguous()
    _lib.check(lib.sg2_split_planes(x.data_ptr(), _lib.ptr(sc), planes.data_ptr(), n, h * wd, c, _lib.stream_ptr(x)), 'sg2_split_planes')
    return planes


def _bwd_prep_planes(gy, y, slope, noise=None, bias=None, d=None, pooled=False, gscale=1.0, signs=None):
    """One pass over (gy, y): gu = gy * lrelu'(y) (slope None: gu = gy, y is not read); returns (planes of gu * d, gb [co],
    gd [n,co] or None).  The per-(sample, channel) sums are reduced in a fixed order (deterministic).
    pooled: gy is the gradient of the 2x2 average pooling that follows the layer (half the resolution); its adjoint
    (broadcast * gscale) is applied while reading.  signs (pooled only): the window signs of y the pooling kernel wrote --
    read instead of y."""
    lib = _lib.load()
    n, co, h, wd = gy.shape
    if pooled:
        h, wd = 2 * h, 2 * wd
    gy = _cl(gy)
    yc = None if ((slope is None and d is None) or signs is not None) else _cl(y)      # y gives the leaky-ReLU sign and, for gd, the accumulator
    assert signs is None or (pooled and d is None and slope is not None)
    planes = torch.empty((2, n, h, wd, co), dtype=torch.bfloat16, device=gy.device)
    gb = torch.empty((co,), dtype=torch.float32, device=gy.device)
    gd = torch.empty((n, co), dtype=torch.float32, device=gy.device) if d is not None else None
    ws = _workspace(lib.sg2_bwd_prep_planes_workspace(n, h * wd, co), gy.device, 'sg2_bwd_prep_planes_workspace')
    f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
    noise, bias, d = f32(noise), (None if bias is None else f32(bias).reshape(-1)), f32(d)
    _lib.check(lib.sg2_bwd_prep_planes(gy.data_ptr(), _lib.ptr(yc), _lib.ptr(noise), _lib.ptr(bias), _lib.ptr(d), planes.data_ptr(),
                                       gb.data_ptr(), _lib.ptr(gd), ws.data_ptr(), _lib.ptr(signs), n, h * wd, co,
                                       float(slope if slope is not None else 1.0), wd if pooled else 0, float(gscale),
                                       _lib.stream_ptr(gy)), 'sg2_bwd_prep_planes')
    return planes, gb, gd


def _conv_planes(xp, w, coef, transpose, accumulate_into=None):
    """y = conv(x, w*coef) (transpose: the data gradient) with x given as pair planes [2][n,h,w,cin]; bf16x3 halo kernel.
    accumulate_into: an existing channels_last result tensor the convolution is ADDED to (returned)."""
    lib = _lib.load()
    co, ci, k, _ = w.shape
    cin, cout = (co, ci) if transpose else (ci, co)
    _, n, h, wd, cx = xp.shape
    assert cx == cin
    packed = _pack(w, coef, transpose, IMPL_HALO)
    if accumulate_into is not None:
        y = accumulate_into
        assert tuple(y.shape) == (n, cout, h, wd) and y.stride() == (h * wd * cout, 1, wd * cout, cout)
    else:
        y = torch.empty_strided((n, cout, h, wd), (h * wd * cout, 1, wd * cout, cout), dtype=torch.float32, device=xp.device)
    with _timed('dgrad' if transpose else 'fwd', 2.0 * n * h * wd * cin * cout * k * k, f'{cin}->{cout} k{k} @{h}x{wd} n{n} planes'):
        _lib.check(lib.sg2_conv2d_fwd_planes(xp.data_ptr(), packed.data_ptr(), y.data_ptr(), _lib.strides4(y), n, h, wd, cin, cout, k,
                                             None, None, 1, 0.0, 1.0, 1 if accumulate_into is not None else 0, _lib.stream_ptr(xp)),
                   'sg2_conv2d_fwd_planes')
    return y


def _wgrad_planes(xp, gyp, k, coef):
    """dw[co,ci,k,k] = coef * sum gy (x) x, both operands as pair planes; deterministic."""
    lib = _lib.load()
    _, n, h, wd, ci = xp.shape
    co = gyp.shape[4]
    dw = torch.empty((co, ci, k, k), dtype=torch.float32, device=xp.device)
    ws = _workspace(lib.sg2_conv2d_wgrad_planes_workspace(n, h, wd, ci, co, k), xp.device, 'sg2_conv2d_wgrad_planes_workspace')
    with _timed('wgrad', 2.0 * n * h * wd * ci * co * k * k, f'{ci}->{co} k{k} @{h}x{wd} n{n} planes'):
        _lib.check(lib.sg2_conv2d_wgrad_planes(xp.data_ptr(), gyp.data_ptr(), dw.data_ptr(), ws.data_ptr(), n, h, wd, ci, co, k,
                                               float(coef), 0, _lib.stream_ptr(xp)), 'sg2_conv2d_wgrad_planes')
    return dw


# ----------------------------------------------------------------------------------------------
# The closed family (any-order autograd).  `coef` is the ELR constant folded into the weight.

_skip_weight_grads = False


class skip_weight_grads:
    """Inside this context the backward of every convolution node returns None for its weight: ``torch.autograd.grad(outputs,
    inputs=image, create_graph=True)`` -- the first-order pass of R1 / path length (nnutils/loss/penalty.py:11-26) -- asks for
    the gradient of the IMAGE only, but a custom Function cannot see that and would launch (and record for double backward)
    every weight-gradient convolution of the discriminator just to have it discarded; ATen's native convolution skips them
    through `task_should_compute_output`.  The second-order terms still reach the weights through Conv2dTransposeFn.backward.
    A process-wide flag (autograd runs CUDA nodes on its own thread), like conv2d_gradfix.weight_gradients_disabled."""

    def __enter__(self):
        global _skip_weight_grads
        self._prev, _skip_weight_grads = _skip_weight_grads, True
        return self

    def __exit__(self, *exc):
        global _skip_weight_grads
        _skip_weight_grads = self._prev


def _need_gw(flag):
    return flag and not _skip_weight_grads


class Conv2dFn(torch.autograd.Function):
    """y = conv2d(x, w * coef), stride 1, same padding."""

    @staticmethod
    @amp_fwd
    def forward(ctx, x, w, coef, precise=True):
        ctx.coef = coef
        ctx.save_for_backward(x, w)
        return _conv_raw(x, w, coef, False, precise=precise)

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = Conv2dTransposeFn.apply(gy, w, ctx.coef)
        if _need_gw(ctx.needs_input_grad[1]):
            gw = Conv2dWgradFn.apply(x, gy, w.shape[2], ctx.coef)
        return gx, gw, None, None


class Conv2dTransposeFn(torch.autograd.Function):
    """gx = conv_transpose2d(gy, w * coef) (the data gradient of Conv2dFn)."""

    @staticmethod
    @amp_fwd
    def forward(ctx, gy, w, coef):
        ctx.coef = coef
        ctx.save_for_backward(gy, w)
        return _conv_raw(gy, w, coef, True)

    @staticmethod
    @amp_bwd
    def backward(ctx, g):
        gy, w = ctx.saved_tensors
        ggy = gw = None
        if ctx.needs_input_grad[0]:
            ggy = Conv2dFn.apply(g, w, ctx.coef, False)       # a gradient quantity: bf16x3 is enough
        if _need_gw(ctx.needs_input_grad[1]):
            gw = Conv2dWgradFn.apply(g, gy, w.shape[2], ctx.coef)
        return ggy, gw, None


class Conv2dWgradFn(torch.autograd.Function):
    """dw[co,ci,k,k] = coef * sum_{n,h,w} gy (x) x (the weight gradient of Conv2dFn)."""

    @staticmethod
    @amp_fwd
    def forward(ctx, x, gy, k, coef):
        ctx.coef = coef
        ctx.save_for_backward(x, gy)
        return _wgrad_raw(x, gy, k, coef)

    @staticmethod
    @amp_bwd
    def backward(ctx, gdw):
        x, gy = ctx.saved_tensors
        gx = ggy = None
        if ctx.needs_input_grad[0]:
            gx = Conv2dTransposeFn.apply(gy, gdw, ctx.coef)
        if ctx.needs_input_grad[1]:
            ggy = Conv2dFn.apply(x, gdw, ctx.coef, False)
        return gx, ggy, None, None


def conv2d(x, w, coef: float = 1.0):
    """Drop-in for ``conv2d_gradfix.conv2d(x, w*coef, padding=k//2)`` (stride 1, odd k)."""
    return Conv2dFn.apply(x, w, float(coef), True)


# ----------------------------------------------------------------------------------------------
# Fused conv + bias + leaky-ReLU (discriminator layers).  Forward is ONE kernel; the backward is
# built from the closed family + the bias_act gradient op, so it stays twice differentiable.

class ConvBiasActFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, w, b, coef, slope, gain=1.0):
        y = _conv_raw(x, w, coef, False, bias=b, slope=slope, gain=gain)
        ctx.coef, ctx.slope, ctx.gain = coef, slope, gain
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    @amp_bwd
    def backward(ctx, gy):
        from .bias_act import act_grad
        x, w, y = ctx.saved_tensors
        co = w.shape[0]
        gb = None
        n_, ci_, h_, wd_ = x.shape
        if (not torch.is_grad_enabled() and ci_ <= 4 and w.shape[2] == 1 and ctx.slope is not None and co in (4, 8, 16, 32, 64)
                and gy.dtype == torch.float32):
            # thin input (from_rgb): leaky-ReLU gradient, bias / weight gradient and (when asked for) the image gradient in ONE pass
            # over (gy, y) -- csrc/thin_bwd.cu
            lib = _lib.load()
            gyc, yc, xc = _cl(gy), _cl(y), _cl(x)
            need_gx, need_gw, need_gb = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
            gx = torch.empty_like(xc) if need_gx else None
            gw = torch.empty_like(w) if need_gw else None
            gb = torch.empty((co,), dtype=torch.float32, device=gy.device) if need_gb else None
            ws = _workspace(lib.sg2_thin_in_bwd_workspace(n_, h_ * wd_, ci_, co), gy.device, 'sg2_thin_in_bwd_workspace')
            _lib.check(lib.sg2_thin_in_bwd(gyc.data_ptr(), yc.data_ptr(), xc.data_ptr(), w.detach().contiguous().data_ptr(), _lib.ptr(gx),
                                           _lib.ptr(gw), _lib.ptr(gb), ws.data_ptr(), n_, h_ * wd_, ci_, co, float(ctx.slope), float(ctx.gain),
                                           float(ctx.coef), _lib.stream_ptr(gy)), 'sg2_thin_in_bwd')
            return gx, gw, gb, None, None, None
        if ctx.gain != 1.0:
            gy = gy * ctx.gain               # y = gain * lrelu(t), gain > 0: sign(y) = sign(t), d y / d t = gain * lrelu'(y)
        if torch.is_grad_enabled() or co % 4 != 0:
            # create_graph (R1): every piece must stay differentiable
            gu = act_grad(gy, y, ctx.slope) if ctx.slope is not None else gy
            if ctx.needs_input_grad[2]:
                gb = gu.sum((0, 2, 3))
        else:
            n, ci, h, wd = x.shape
            k = w.shape[2]
            need_gx, need_gw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            use_planes = ((need_gx or need_gw) and (not need_gx or _planes_ok(n, h, wd, co, ci, k, False))
                          and (not need_gw or _planes_ok(n, h, wd, ci, co, k, True)))
            if use_planes:
                # first-order fast path: ONE pass over (gy, y) writes gu = gy * lrelu'(y) as bf16 pair planes (+ the bias
                # gradient); data and weight gradient take them by TMA, no fp32 -> bf16 transform inside the kernels
                gup, gb, _ = _bwd_prep_planes(gy, y, ctx.slope)
                gx = _conv_planes(gup, w, ctx.coef, True) if need_gx else None
                gw = _wgrad_planes(_split_planes(x), gup, k, ctx.coef) if need_gw else None
                return gx, gw, (gb if ctx.needs_input_grad[2] else None), None, None, None
            # first-order backward: leaky-ReLU mask and the bias-gradient reduction in ONE pass over (gy, y)
            lib = _lib.load()
            gyc = _cl(gy)
            part = torch.empty((n, co), dtype=torch.float32, device=gy.device)
            ws = _workspace(lib.sg2_reduce_hw_workspace(n, h * wd, co), gy.device, 'sg2_reduce_hw_workspace')
            if ctx.slope is not None:
                gu = _empty_cl(n, co, h, wd, gy)
                _lib.check(lib.sg2_modconv_bwd_prep(gyc.data_ptr(), _cl(y).data_ptr(), None, None, None, gu.data_ptr(),
                                                    part.data_ptr(), None, n, h * wd, co, float(ctx.slope),
                                                    ws.data_ptr(), _lib.stream_ptr(gy)), 'sg2_modconv_bwd_prep')
            else:
                gu = gyc
                _lib.check(lib.sg2_reduce_hw(gyc.data_ptr(), None, part.data_ptr(), n, h * wd, co, ws.data_ptr(),
                                             _lib.stream_ptr(gy)), 'sg2_reduce_hw')
            gb = part.sum(0)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = Conv2dTransposeFn.apply(gu, w, ctx.coef)
        if _need_gw(ctx.needs_input_grad[1]):
            gw = Conv2dWgradFn.apply(x, gu, w.shape[2], ctx.coef)
        return gx, gw, (gb if ctx.needs_input_grad[2] else None), None, None, None


def conv2d_bias_act(x, w, b, coef: float = 1.0, slope: float | None = 0.2, gain: float = 1.0):
    """gain * lrelu_slope(conv2d(x * coef, w) + b): Conv2d('elr') + LeakyReLU (model.py:50-53, 191-193); with gain = the
    bias_act gain of the StyleGAN3-style ConvAct (implementations/StyleGAN3/model.py:411-416).
    slope=None -> no activation (the DBlock skip conv, model.py:201)."""
    assert gain > 0
    return ConvBiasActFn.apply(x, w, b, float(coef), slope, float(gain))


# ----------------------------------------------------------------------------------------------
# The whole residual discriminator block (DBlock.forward, implementations/StyleGAN2/model.py:204-212) as ONE autograd node:
#   h1 = lrelu(conv3x3(x) + b1);  h2 = lrelu(conv3x3(h1) + b2);  t = conv1x1(x) + bs;  out = alpha * (avgpool2(h2) + avgpool2(t))
# What the single node buys: the skip branch runs at the pooled resolution (see forward), and in the first-order backward the
# pooling adjoint is folded into the leaky-ReLU-gradient pass that follows it (the full-resolution gradient of the pooling
# input is never written or read) and the two data gradients that meet at x accumulate in a kernel epilogue instead of a
# separate add.  Under create_graph (R1) the backward is composed from the differentiable families, exactly like
# ConvBiasActFn.

def _composed_conv_backward(x, w, y, gy, coef, slope, need_gx, need_gw, need_gb):
    from .bias_act import act_grad
    gu = act_grad(gy, y, slope) if slope is not None else gy
    gx = Conv2dTransposeFn.apply(gu, w, coef) if need_gx else None
    gw = Conv2dWgradFn.apply(x, gu, w.shape[2], coef) if need_gw else None
    gb = gu.sum((0, 2, 3)) if need_gb else None
    return gx, gw, gb


class DBlockFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, w1, b1, w2, b2, ws, bs, coef1, coef2, coefs, slope, alpha):
        from .resample import _avgpool
        n, ci, h, wd = x.shape
        h1 = _conv_raw(x, w1, coef1, False, bias=b1, slope=slope)
        h2 = _conv_raw(h1, w2, coef2, False, bias=b2, slope=slope)
        # The skip branch down(skip(x)) (model.py:207-211) is evaluated as skip(down(x)): a 1x1 convolution (+ bias) commutes with
        # the 2x2 average that follows it, so the branch runs on a quarter of the pixels -- forward, data gradient and weight
        # gradient -- and its full-resolution output t is never written.  Same function, different rounding order (~1e-7).
        pooled_skip = h % 2 == 0 and wd % 2 == 0 and ci % 4 == 0
        if pooled_skip:
            xq = _avgpool(x, None, 1.0, False)
            tq = _conv_raw(xq, ws, coefs, False, bias=bs)
            # the pooling kernel reads h2 anyway: it also records the leaky-ReLU signs of every window (2 bytes per 16 elements), so
            # the backward pass that follows the pooling adjoint reads those instead of h2 (64 bytes per window)
            out, signs2 = _avgpool(h2, tq, alpha, False, t_pooled=True, want_signs=True)
        else:
            xq = signs2 = torch.empty(0, device=x.device)
            out = _avgpool(h2, _conv_raw(x, ws, coefs, False, bias=bs), alpha, False)
        ctx.cfg = (coef1, coef2, coefs, slope, alpha, pooled_skip)
        ctx.save_for_backward(x, w1, w2, ws, h1, h2, xq, signs2)
        return out

    @staticmethod
    @amp_bwd
    def backward(ctx, g):
        from .resample import AvgPool2AdjFn, _avgpool
        x, w1, w2, ws, h1, h2, xq, signs2 = ctx.saved_tensors
        coef1, coef2, coefs, slope, alpha, pooled_skip = ctx.cfg
        need = ctx.needs_input_grad
        n, ci, h, wd = x.shape
        co = w1.shape[0]
        fast = (not torch.is_grad_enabled() and pooled_skip and co % 4 == 0
                and _planes_ok(n, h, wd, co, ci, 3, False) and _planes_ok(n, h, wd, co, co, 3, False)
                and _planes_ok(n, h, wd, ci, co, 3, True) and _planes_ok(n, h, wd, co, co, 3, True))
        # the skip branch lives at the pooled resolution, which the planes kernels may not take (4 x 4 images): fp32-operand kernels then
        skip_planes = _planes_ok(n, h // 2, wd // 2, co, ci, 1, False) and _planes_ok(n, h // 2, wd // 2, ci, co, 1, True)
        if not fast:
            # composed from the differentiable families (create_graph, odd shapes): the gradient of the same function, written in
            # the reference's order -- skip at full resolution, then the pooling adjoint
            gf = AvgPool2AdjFn.apply(g, alpha)
            gh1, gw2, gb2 = _composed_conv_backward(h1, w2, h2, gf, coef2, slope, True, _need_gw(need[3]), need[4])
            gx1, gw1, gb1 = _composed_conv_backward(x, w1, h1, gh1, coef1, slope, need[0], _need_gw(need[1]), need[2])
            gx2, gws, gbs = _composed_conv_backward(x, ws, None, gf, coefs, None, need[0], _need_gw(need[5]), need[6])
            gx = gx1 + gx2 if need[0] else None
            return gx, gw1, gb1, gw2, gb2, gws, gbs, None, None, None, None, None
        gu2p, gb2, _ = _bwd_prep_planes(g, h2, slope, pooled=True, gscale=0.25 * alpha,       # d out / d h2, masked by lrelu'(h2)
                                        signs=signs2 if signs2.numel() else None)
        if skip_planes:
            gtp, gbs, _ = _bwd_prep_planes(g, None, None, gscale=alpha)                     # d out / d tq, at the pooled resolution
        else:
            gt = g * alpha
            gbs = gt.sum((0, 2, 3))
        gh1 = _conv_planes(gu2p, w2, coef2, True)
        gu1p, gb1, _ = _bwd_prep_planes(gh1, h1, slope)
        gx = None
        if need[0]:
            # skip branch first: its data gradient at the pooled resolution, spread over the 2x2 windows (pooling adjoint) into
            # gx; the main branch's data gradient then accumulates on top in the epilogue of its kernel
            gxq = _conv_planes(gtp, ws, coefs, True) if skip_planes else _conv_raw(gt, ws, coefs, True)
            gx = _conv_planes(gu1p, w1, coef1, True, accumulate_into=_avgpool(gxq, None, 1.0, True))
        gw1 = _wgrad_planes(_split_planes(x), gu1p, 3, coef1) if need[1] else None
        gws = None
        if need[5]:
            gws = _wgrad_planes(_split_planes(xq), gtp, 1, coefs) if skip_planes else _wgrad_raw(xq, gt, 1, coefs)
        gw2 = _wgrad_planes(_split_planes(h1), gu2p, 3, coef2) if need[3] else None
        return (gx, gw1, gb1 if need[2] else None, gw2, gb2 if need[4] else None, gws, gbs if need[6] else None,
                None, None, None, None, None)


def dblock(x, w1, b1, w2, b2, ws, bs, coef1, coef2, coefs, slope=0.2, alpha=0.7071067811865476):
    """DBlock.forward (model.py:204-212) for the standard block (two 3x3 convolutions, 1x1 skip, average pooling)."""
    return DBlockFn.apply(x, w1, b1, w2, b2, ws, bs, float(coef1), float(coef2), float(coefs), float(slope), float(alpha))


# ----------------------------------------------------------------------------------------------
# Modulated convolution (generator).  The per-sample weight tensor [B, Co, Ci, k, k] of the reference
# (model.py:115-120) is never materialised.  ModConvFn is the fused first-order path of every ordinary step;
# path-length steps run the any-order composition below (same arithmetic, separate kernels).

class ModConvFn(torch.autograd.Function):
    @staticmethod
    @amp_fwd
    def forward(ctx, x, w, s, d, b, noise, coef, slope, out_nchw):
        y = _conv_raw(x, w, coef, False, in_scale=s, out_scale=d, bias=b, noise=noise, slope=slope, out_nchw=out_nchw)
        ctx.coef, ctx.slope, ctx.out_nchw = coef, slope, out_nchw
        ctx.save_for_backward(x, w, s, d if d is not None else torch.empty(0, device=x.device),
                              b if b is not None else torch.empty(0, device=x.device),
                              noise if noise is not None else torch.empty(0, device=x.device), y)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    @amp_bwd
    def backward(ctx, gy):
        lib = _lib.load()
        x, w, s, d, b, noise, y = ctx.saved_tensors
        d = d if d.numel() else None
        b = b if b.numel() else None
        noise = noise if noise.numel() else None
        n, ci, h, wd = x.shape
        co, k = w.shape[0], w.shape[2]
        st = _lib.stream_ptr(x)
        x = _cl(x)
        need_gw = ctx.needs_input_grad[1]
        use_planes = (co % 4 == 0 and _planes_ok(n, h, wd, co, ci, k, False) and (not need_gw or _planes_ok(n, h, wd, ci, co, k, True)))
        g_acc_p = None
        if use_planes:
            g_acc_p, gb, gd = _bwd_prep_planes(gy, y, ctx.slope, noise=noise, bias=b, d=d)
        elif co % 4 == 0:
            gy, yc = _cl(gy), _cl(y)
            g_acc = _empty_cl(n, co, h, wd, x)
            gb_part = torch.empty((n, co), dtype=torch.float32, device=x.device)
            gd = torch.empty((n, co), dtype=torch.float32, device=x.device) if d is not None else None
            bflat = None if b is None else b.detach().reshape(-1).contiguous()
            ws = _workspace(lib.sg2_reduce_hw_workspace(n, h * wd, co), x.device, 'sg2_reduce_hw_workspace')
            _lib.check(lib.sg2_modconv_bwd_prep(
                gy.data_ptr(), yc.data_ptr(), _lib.ptr(noise), _lib.ptr(bflat), _lib.ptr(d), g_acc.data_ptr(),
                gb_part.data_ptr(), _lib.ptr(gd), n, h * wd, co,
                float(ctx.slope if ctx.slope is not None else 1.0), ws.data_ptr(), st), 'sg2_modconv_bwd_prep')
            gb = gb_part.sum(0)
        else:
            # output channel counts the vectorised prologue does not take (ToRGB's Co = 3: tiny tensors; the odd widths of
            # the StyleGAN3 generator): the same arithmetic as sg2_modconv_bwd_prep in tensor ops
            gu = gy if ctx.slope is None else gy * torch.where(y > 0, 1.0, float(ctx.slope))
            gb = gu.sum((0, 2, 3))
            gd = None
            if d is not None:
                u = y if ctx.slope is None else torch.where(y > 0, y, y / float(ctx.slope))
                if b is not None:
                    u = u - b.reshape(1, -1, 1, 1)
                if noise is not None:
                    u = u - noise
                gd = (gu * u).sum((2, 3)) / d
                gu = gu * d[:, :, None, None]
            g_acc = gu
        gx = gw = gs = None
        # data gradient w.r.t. the modulated input, then gs = sum_hw g_xs * x and gx = g_xs * s in one pass
        g_xs = _conv_planes(g_acc_p, w, ctx.coef, True) if use_planes else _conv_raw(g_acc, w, ctx.coef, True)
        if ci % 4 == 0:
            gs = torch.empty((n, ci), dtype=torch.float32, device=x.device)
            gx = _empty_cl(n, ci, h, wd, x) if ctx.needs_input_grad[0] else None
            sc = s.detach().contiguous()
            ws = _workspace(lib.sg2_reduce_hw_workspace(n, h * wd, ci), x.device, 'sg2_reduce_hw_workspace')
            _lib.check(lib.sg2_scale_reduce_hw(g_xs.data_ptr(), x.data_ptr(), sc.data_ptr(), _lib.ptr(gx),
                                               gs.data_ptr(), n, h * wd, ci, ws.data_ptr(), st), 'sg2_scale_reduce_hw')
        else:
            gs = (g_xs * x).sum((2, 3))
            gx = g_xs * s[:, :, None, None]
        if need_gw:
            if use_planes:
                gw = _wgrad_planes(_split_planes(x, s), g_acc_p, k, ctx.coef)
            else:
                gw = _wgrad_raw(x, g_acc, k, ctx.coef, in_scale=s)
        if b is not None:
            gb = gb.reshape(b.shape)
        return gx, gw, gs, gd, (gb if b is not None else None), None, None, None, None


class DemodFn(torch.autograd.Function):
    """d[b,o] = rsqrt(coef^2 * sum_i s[b,i]^2 * sum_k w[o,i,k]^2 + eps) (model.py:118-120 on the [B,Co] coefficient), first order
    (path-length steps, which differentiate twice through d, run the tensor expression inside any_order_modconv)."""

    @staticmethod
    @amp_fwd
    def forward(ctx, w, s, coef, eps):
        lib = _lib.load()
        co, ci, k, _ = w.shape
        w, s = w.contiguous(), s.contiguous()
        B = s.shape[0]
        wsq = torch.empty((co, ci), dtype=torch.float32, device=w.device)
        d = torch.empty((B, co), dtype=torch.float32, device=w.device)
        _lib.check(lib.sg2_demod_fwd(w.data_ptr(), s.data_ptr(), wsq.data_ptr(), d.data_ptr(), B, co, ci, k * k, float(coef), float(eps),
                                     _lib.stream_ptr(w)), 'sg2_demod_fwd')
        ctx.coef = coef
        ctx.save_for_backward(w, s, wsq, d)
        return d

    @staticmethod
    @torch.autograd.function.once_differentiable
    @amp_bwd
    def backward(ctx, gd):
        lib = _lib.load()
        w, s, wsq, d = ctx.saved_tensors
        co, ci, k, _ = w.shape
        B = s.shape[0]
        gd = gd.to(torch.float32).contiguous()
        gw = torch.empty_like(w) if ctx.needs_input_grad[0] else None
        gs = torch.empty_like(s) if ctx.needs_input_grad[1] else None
        _lib.check(lib.sg2_demod_bwd(w.data_ptr(), s.data_ptr(), wsq.data_ptr(), d.data_ptr(), gd.data_ptr(), _lib.ptr(gw), _lib.ptr(gs),
                                     B, co, ci, k * k, float(ctx.coef), _lib.stream_ptr(w)), 'sg2_demod_bwd')
        return gw, gs, None, None


_any_order = False


class any_order_modconv:
    """Context manager: inside it ``modulated_conv2d`` is composed from the closed convolution family, the
    twice-differentiable ``bias_act`` and torch broadcasts, so gradients of ANY order exist -- what the
    path-length regulariser needs (second order through the modulated convolution w.r.t. the style,
    implementations/StyleGAN2/utils.py:18-29).  Outside it the fused first-order ``ModConvFn`` runs."""

    def __enter__(self):
        global _any_order
        self._prev, _any_order = _any_order, True
        return self

    def __exit__(self, *exc):
        global _any_order
        _any_order = self._prev


def _modulated_conv2d_any_order(x, w, s, d, bias, noise, coef, slope):
    from .bias_act import bias_act
    acc = Conv2dFn.apply(x * s[:, :, None, None], w, coef, True)
    if d is not None:
        acc = acc * d[:, :, None, None]
    if bias is not None:
        acc = acc + bias
    if noise is not None:
        acc = acc + noise
    if slope is not None:
        acc = bias_act(acc, None, act='lrelu', alpha=slope, gain=1.0)
    return acc


def modulated_conv2d(x, w, s, bias=None, noise=None, demod=True, slope=None, eps=1e-4, out_nchw=False, in_gain=None):
    """ModulatedConv2d.forward (model.py:106-132) (+ InjectNoise :85-88 + LeakyReLU :164 when given).

    x [B,Ci,H,W]; w [Co,Ci,k,k]; s [B,Ci] = affine(style) + 1 (model.py:110); bias [1,Co,1,1];
    noise [B,1,H,W].  The demodulation coefficient d[b,o] = rsqrt(sum_i s^2 * sum_k (w*coef)^2 + eps)
    is a tiny [B,Ci]x[Ci,Co] product kept in torch so its gradients reach `s` and `w` through autograd."""
    co, ci, k, _ = w.shape
    coef = 1.0 / float(ci * k * k) ** 0.5
    d = None
    if demod:
        if (not _any_order and w.is_cuda and w.dtype == torch.float32 and s.dtype == torch.float32 and ci <= 2048 and co <= 2048
                and s.shape[0] <= 256):
            d = DemodFn.apply(w, s, coef, float(eps))          # one launch (two in the backward) instead of ~7 (~12)
        else:
            wsq = w.square().sum((2, 3))                       # [Co,Ci]
            d = torch.rsqrt(torch.matmul(s.square(), wsq.t()) * (coef * coef) + eps)
    if in_gain is not None:
        # StyleGAN3's magnitude-EMA input gain (implementations/StyleGAN3/model.py:62-65): scales the weight per input channel
        # AFTER demodulation, i.e. it rides with the style on the activation tile but stays out of d
        s = s * in_gain
    if _any_order:
        return _modulated_conv2d_any_order(x, w, s, d, bias, noise, coef, slope)
    return ModConvFn.apply(x, w, s, d, bias, noise, coef, slope, out_nchw)
