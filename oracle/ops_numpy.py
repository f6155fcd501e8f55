"""CPU ORACLE (test infrastructure, NOT a product path) -- numpy restatement of the reference ops.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Every function restates the algorithm of a reference function (cited file:line, relative to the
STomoya/animeface checkout) in plain numpy, written from the definition rather than translated.

PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py imports the reference
from /root/reference in the build container and stores seeded input/output vectors in tests/golden/*.npz;
tests/test_oracle_golden.py checks every function below against them on CPU.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------------------------
# upfirdn2d  -- thirdparty/stylegan3_ops/ops/upfirdn2d.py:161-207 (_upfirdn2d_ref); out size upfirdn2d.cpp:29-30

def _pair(v):
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _quad(p):
    if isinstance(p, int):
        return p, p, p, p
    if len(p) == 2:
        return p[0], p[0], p[1], p[1]
    return tuple(p)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1.0):
    """x [N,C,H,W]; f [fh,fw] or [taps] (separable) or None.  Direct evaluation of the definition:
    y[jy,jx] = gain * sum_t U[jy*dy + ty - py0, jx*dx + tx - px0] * wt[ty,tx], U = zero-stuffed x,
    wt = f flipped unless flip_filter (true convolution by default)."""
    x = np.asarray(x)
    acc_t = np.float64 if x.dtype == np.float64 else np.float32
    upx, upy = _pair(up)
    dnx, dny = _pair(down)
    px0, px1, py0, py1 = _quad(padding)
    if f is None:
        f = np.ones((1, 1), np.float32)
    f = np.asarray(f, np.float32)
    if f.ndim == 1:
        # separable: horizontal pass with gain 1 then vertical pass with the gain (upfirdn2d.py:238-239)
        y = upfirdn2d(x, f[None, :], (upx, 1), (dnx, 1), (px0, px1, 0, 0), flip_filter, 1.0)
        return upfirdn2d(y, f[:, None], (1, upy), (1, dny), (0, 0, py0, py1), flip_filter, gain)
    n, c, h, w = x.shape
    fh, fw = f.shape
    # zero-stuffed, padded (negative = crop) canvas
    uh, uw = h * upy, w * upx
    canvas = np.zeros((n, c, uh + max(py0, 0) + max(py1, 0), uw + max(px0, 0) + max(px1, 0)), acc_t)
    canvas[:, :, max(py0, 0):max(py0, 0) + uh:upy, max(px0, 0):max(px0, 0) + uw:upx] = x
    canvas = canvas[:, :, max(-py0, 0):canvas.shape[2] - max(-py1, 0), max(-px0, 0):canvas.shape[3] - max(-px1, 0)]
    ch, cw = canvas.shape[2], canvas.shape[3]
    oh_full, ow_full = ch - fh + 1, cw - fw + 1
    assert oh_full >= 1 and ow_full >= 1
    wt = f if flip_filter else f[::-1, ::-1]
    out = np.zeros((n, c, oh_full, ow_full), acc_t)
    for ty in range(fh):
        for tx in range(fw):
            out += canvas[:, :, ty:ty + oh_full, tx:tx + ow_full] * acc_t(wt[ty, tx])
    out = out * acc_t(gain)
    return out[:, :, ::dny, ::dnx].astype(x.dtype)


def filter2d(x, f, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:271-303"""
    px0, px1, py0, py1 = _quad(padding)
    fh, fw = (f.shape[0], f.shape[-1]) if f is not None else (1, 1)
    return upfirdn2d(x, f, padding=(px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2),
                     flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:307-342"""
    ux, uy = _pair(up)
    px0, px1, py0, py1 = _quad(padding)
    fh, fw = f.shape[0], f.shape[-1]
    p = (px0 + (fw + ux - 1) // 2, px1 + (fw - ux) // 2, py0 + (fh + uy - 1) // 2, py1 + (fh - uy) // 2)
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * ux * uy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:346-381"""
    dx, dy = _pair(down)
    px0, px1, py0, py1 = _quad(padding)
    fh, fw = f.shape[0], f.shape[-1]
    p = (px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2)
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


def setup_filter(f, normalize=True, flip_filter=False, gain=1.0, separable=None):
    """upfirdn2d.py:64-108"""
    f = np.asarray(1 if f is None else f, np.float32)
    if f.ndim == 0:
        f = f[None]
    sep = (f.ndim == 1 and f.size >= 8) if separable is None else separable
    if f.ndim == 1 and not sep:
        f = np.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f[::-1].copy() if f.ndim == 1 else f[::-1, ::-1].copy()
    return (f * gain ** (f.ndim / 2)).astype(np.float32)


# --------------------------------------------------------------------------------------------------------
# StyleGAN2 resampling -- implementations/StyleGAN2/model.py:56-63, 138-149

def bilinear_up2x(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False): source coordinate of output j is
    (j + 0.5)/2 - 0.5, clamped to the image (replicate border)."""
    x = np.asarray(x, np.float32)

    def axis(a, ax):
        n = a.shape[ax]
        j = np.arange(2 * n)
        src = np.clip((j + 0.5) / 2 - 0.5, 0, None)
        i0 = np.minimum(np.floor(src).astype(int), n - 1)
        i1 = np.minimum(i0 + 1, n - 1)
        lam = (src - i0).astype(np.float32)
        shape = [1] * a.ndim
        shape[ax] = 2 * n
        lam = lam.reshape(shape)
        return np.take(a, i0, ax) * (1 - lam) + np.take(a, i1, ax) * lam

    return axis(axis(x, 2), 3)


def blur3x3(x):
    """Blur2d: depthwise [1,2,1]x[1,2,1]/16, zero padding 1 (model.py:138-149)."""
    x = np.asarray(x, np.float32)
    k = np.array([1., 2., 1.], np.float32)
    k2 = np.outer(k, k) / 16
    p = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    out = np.zeros_like(x)
    h, w = x.shape[2:]
    for a in range(3):
        for b in range(3):
            out += p[:, :, a:a + h, b:b + w] * k2[a, b]
    return out


def avgpool2(x):
    """nn.AvgPool2d(2) (model.py:61-63)."""
    x = np.asarray(x, np.float32)
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).mean((3, 5))


# --------------------------------------------------------------------------------------------------------
# bias_act -- thirdparty/stylegan3_ops/ops/bias_act.py:86-115 (_bias_act_ref), table :16-26

_SQRT2 = float(np.sqrt(2))
ACT_DEFAULTS = {  # name: (def_alpha, def_gain)
    'linear': (0, 1), 'relu': (0, _SQRT2), 'lrelu': (0.2, _SQRT2), 'tanh': (0, 1), 'sigmoid': (0, 1),
    'elu': (0, 1), 'selu': (0, 1), 'softplus': (0, 1), 'swish': (0, _SQRT2),
}


def _act(name, x, alpha):
    if name == 'linear':
        return x
    if name == 'relu':
        return np.maximum(x, 0)
    if name == 'lrelu':
        return np.where(x > 0, x, x * alpha)
    if name == 'tanh':
        return np.tanh(x)
    if name == 'sigmoid':
        return 1 / (1 + np.exp(-x))
    if name == 'elu':
        return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    if name == 'selu':
        sc, al = 1.0507009873554804934193349852946, 1.6732632423543772848170429916717
        return sc * np.where(x > 0, x, al * np.expm1(np.minimum(x, 0)))
    if name == 'softplus':
        return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))
    if name == 'swish':
        return x / (1 + np.exp(-x))
    raise KeyError(name)


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    x = np.asarray(x)
    da, dg = ACT_DEFAULTS[act]
    alpha = float(da if alpha is None else alpha)
    gain = float(dg if gain is None else gain)
    if b is not None:
        shape = [1] * x.ndim
        shape[dim] = -1
        x = x + np.asarray(b).reshape(shape)
    y = _act(act, x, alpha)
    if gain != 1:
        y = y * gain
    if clamp is not None and clamp >= 0:
        y = np.clip(y, -clamp, clamp)
    return y.astype(x.dtype)


# --------------------------------------------------------------------------------------------------------
# minibatch stddev -- implementations/StyleGAN2/model.py:215-236

def minibatch_stddev(x, group_size, eps=1e-4):
    x = np.asarray(x, np.float32)
    b, c, h, w = x.shape
    g = group_size if b % group_size == 0 else b
    y = x.reshape(g, b // g, c, h, w)
    y = y - y.mean(0, keepdims=True)
    y = np.sqrt((y * y).mean(0) + eps)               # [M,C,H,W]
    y = y.mean((1, 2, 3), keepdims=True)             # [M,1,1,1]
    y = np.tile(y, (g, 1, h, w))                     # [B,1,H,W]  sample i -> column i % M
    return np.concatenate([x, y], 1)


# --------------------------------------------------------------------------------------------------------
# convolution -- F.conv2d stride 1 same padding; ModulatedConv2d.forward model.py:106-132; ELR model.py:29-37

def conv2d_same(x, w):
    """Cross-correlation, stride 1, zero padding (k-1)//2.  x [N,Ci,H,W], w [Co,Ci,k,k]; float64 accumulate."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, ci, h, wd = x.shape
    co, _, k, _ = w.shape
    pad = (k - 1) // 2
    p = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    out = np.zeros((n, co, h, wd), np.float64)
    for a in range(k):
        for b in range(k):
            out += np.einsum('nchw,oc->nohw', p[:, :, a:a + h, b:b + wd], w[:, :, a, b])
    return out


def modulated_conv2d(x, w, s, bias=None, demod=True, eps=1e-4):
    """Per-sample weights Wm[b] = W * s[b] * coef, optional demodulation, grouped conv, + bias
    (model.py:110-132; `s` is affine(style) + 1)."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    s = np.asarray(s, np.float64)
    co, ci, k, _ = w.shape
    coef = 1.0 / np.sqrt(ci * k * k)
    out = []
    for b in range(x.shape[0]):
        wm = w * s[b][None, :, None, None] * coef
        if demod:
            wm = wm / np.sqrt((wm ** 2).sum((1, 2, 3), keepdims=True) + eps)
        out.append(conv2d_same(x[b:b + 1], wm))
    y = np.concatenate(out, 0)
    if bias is not None:
        y = y + np.asarray(bias, np.float64).reshape(1, -1, 1, 1)
    return y


def elr_conv2d(x, w, b=None):
    """ELR(nn.Conv2d): conv(x * coef, w) + b with coef = 1/sqrt(fan_in) (model.py:29-37, 50-53)."""
    coef = 1.0 / np.sqrt(np.prod(w.shape[1:]))
    y = conv2d_same(np.asarray(x, np.float64) * coef, w)
    if b is not None:
        y = y + np.asarray(b, np.float64).reshape(1, -1, 1, 1)
    return y


# --------------------------------------------------------------------------------------------------------
# Adam / EMA -- torch.optim.Adam as configured at implementations/StyleGAN2/utils.py:208-221; nnutils/training.py:23-40

def adam_step(p, g, m, v, step, lr, b1, b2, eps=1e-8):
    """One torch.optim.Adam update (no amsgrad / weight decay); returns new (p, m, v). `step` is 1-based."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p = p - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
    return p, m, v


def ema_step(ema, p, decay=0.999):
    return ema * decay + p * (1 - decay)
