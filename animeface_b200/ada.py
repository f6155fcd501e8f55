"""Adaptive discriminator augmentation on the libsg2b200 ops (BASELINE config 4).

Drop-in for ``thirdparty/ada/augment.py`` (``AugmentPipe`` :115-427) and ``nnutils/ada.py`` (``ADA`` :5-36): same constructor
arguments, buffers (``p``, ``Hz_geom``, ``Hz_fbank``, ``signsum``), random-draw order and arithmetic.  What differs is the
execution of the image-sized work:
  * geometry (:270-299): reflect padding, x2 up-sampling with the 12-tap sym6 low-pass, the inverse-warp resampling and the
    x2 down-sampling run on ``ops.grid_sample.reflect_pad`` / ``ops.upfirdn2d`` / ``ops.grid_sample.affine_grid_sample`` --
    the sampling grid ([B, 2H', 2W', 2] floats in the reference) is never materialised, and every op is differentiable to
    any order (the reference needs ``grid_sample_gradfix`` for that);
  * colour (:352-361): one pass applying the per-sample 3x4 matrix (``ops.grid_sample.color_affine``).
The per-sample 3x3 / 4x4 transform matrices are composed with ordinary tensor algebra on [B, 3, 3] / [B, 4, 4] tensors --
control logic, a few hundred bytes.  Draws go through ``animeface_b200.rng`` so tests can replay the reference's.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import rng
from .ops import upfirdn2d
from .ops.grid_sample import affine_grid_sample, color_affine, reflect_pad

# Low-pass coefficients of the symlet wavelets used by the pipeline (standard values; the reference tabulates them at
# thirdparty/ada/augment.py:19-36): sym6 for the geometric resampling, sym2 for the image-space filter bank.
SYM2 = [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025]
SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466, 0.787641141030194,
        0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148]


# ---- batched homogeneous matrices ---------------------------------------------------------------------------------
def _const(v, like):
    return torch.full_like(like, float(v))


def _mat(rows, like):
    """rows of scalars / [B]-shaped tensors -> [B, r, c]."""
    cols = [torch.stack([e if isinstance(e, torch.Tensor) else _const(e, like) for e in row], dim=-1) for row in rows]
    return torch.stack(cols, dim=-2)


def _translate2(tx, ty):
    ref = tx if isinstance(tx, torch.Tensor) else ty
    return _mat([[1, 0, tx], [0, 1, ty], [0, 0, 1]], ref)


def _scale2(sx, sy):
    ref = sx if isinstance(sx, torch.Tensor) else sy
    return _mat([[sx, 0, 0], [0, sy, 0], [0, 0, 1]], ref)


def _rotate2(theta):
    c, s = torch.cos(theta), torch.sin(theta)
    return _mat([[c, -s, 0], [s, c, 0], [0, 0, 1]], theta)


def _filter_bank():
    """Four band-pass filters of a 3-level sym2 decomposition, H_i(z) (reference :173-183): start from a unit impulse in
    band 0; each level zero-stuffs every filter by 2, smooths with H(z)H(1/z)/2 and adds the level's high-pass
    H(-z)H(-1/z)/2, centred, to the next band."""
    import scipy.signal
    lo = np.asarray(SYM2)
    hi = lo * ((-1) ** np.arange(lo.size))
    lo2 = np.convolve(lo, lo[::-1]) / 2
    hi2 = np.convolve(hi, hi[::-1]) / 2
    bank = np.eye(4, 1)
    for i in range(1, bank.shape[0]):
        bank = np.dstack([bank, np.zeros_like(bank)]).reshape(bank.shape[0], -1)[:, :-1]
        bank = scipy.signal.convolve(bank, [lo2])
        mid = bank.shape[1]
        bank[i, (mid - hi2.size) // 2:(mid + hi2.size) // 2] += hi2
    return torch.as_tensor(bank, dtype=torch.float32)


class AugmentPipe(torch.nn.Module):
    """All augmentations are off unless their probability multiplier is set (reference :115-183 for the arguments)."""

    def __init__(self, xflip=0, rotate90=0, xint=0, xint_max=0.125,
                 scale=0, rotate=0, aniso=0, xfrac=0, scale_std=0.2, rotate_max=1, aniso_std=0.2, xfrac_std=0.125,
                 brightness=0, contrast=0, lumaflip=0, hue=0, saturation=0, brightness_std=0.2, contrast_std=0.5, hue_max=1, saturation_std=1,
                 imgfilter=0, imgfilter_bands=[1, 1, 1, 1], imgfilter_std=1,
                 noise=0, cutout=0, noise_std=0.1, cutout_size=0.5):
        super().__init__()
        self.register_buffer('p', torch.ones([]))       # overall multiplier of every augmentation probability
        for name, val in dict(xflip=xflip, rotate90=rotate90, xint=xint, xint_max=xint_max, scale=scale, rotate=rotate, aniso=aniso,
                              xfrac=xfrac, scale_std=scale_std, rotate_max=rotate_max, aniso_std=aniso_std, xfrac_std=xfrac_std,
                              brightness=brightness, contrast=contrast, lumaflip=lumaflip, hue=hue, saturation=saturation,
                              brightness_std=brightness_std, contrast_std=contrast_std, hue_max=hue_max, saturation_std=saturation_std,
                              imgfilter=imgfilter, imgfilter_std=imgfilter_std, noise=noise, cutout=cutout, noise_std=noise_std,
                              cutout_size=cutout_size).items():
            setattr(self, name, float(val))
        self.imgfilter_bands = list(imgfilter_bands)
        self.register_buffer('Hz_geom', upfirdn2d.setup_filter(SYM6))
        self.register_buffer('Hz_fbank', _filter_bank())

    # ---- parameter sampling: the reference's draws, in the reference's order ----------------------------------------
    def _gate(self, shape, prob, value, neutral, dev):
        """value where a fresh uniform draw < prob * p, else the neutral element."""
        return torch.where(rng.rand(*shape, device=dev) < prob * self.p, value, torch.full_like(value, neutral))

    def _geometry(self, B, width, height, dev, pct):
        G = None                                    # inverse transform: G @ pixel_out -> pixel_in
        mul = lambda M: M if G is None else G @ M
        if self.xflip > 0:
            i = self._gate([B], self.xflip, torch.floor(rng.rand(B, device=dev) * 2), 0, dev)
            if pct is not None:
                i = torch.full_like(i, float(torch.floor(pct * 2)))
            G = mul(_scale2(1 / (1 - 2 * i), _const(1, i)))
        if self.rotate90 > 0:
            i = self._gate([B], self.rotate90, torch.floor(rng.rand(B, device=dev) * 4), 0, dev)
            if pct is not None:
                i = torch.full_like(i, float(torch.floor(pct * 4)))
            G = mul(_rotate2(math.pi / 2 * i))
        if self.xint > 0:
            t = self._gate([B, 1], self.xint, (rng.rand(B, 2, device=dev) * 2 - 1) * self.xint_max, 0, dev)
            if pct is not None:
                t = torch.full_like(t, float((pct * 2 - 1) * self.xint_max))
            G = mul(_translate2(-torch.round(t[:, 0] * width), -torch.round(t[:, 1] * height)))
        if self.scale > 0:
            s = self._gate([B], self.scale, torch.exp2(rng.randn(B, device=dev) * self.scale_std), 1, dev)
            if pct is not None:
                s = torch.full_like(s, float(torch.exp2(torch.erfinv(pct * 2 - 1) * self.scale_std)))
            G = mul(_scale2(1 / s, 1 / s))
        p_rot = 1 - torch.sqrt((1 - self.rotate * self.p).clamp(0, 1))        # P(pre or post rotation) = rotate * p
        for stage in ('pre', 'aniso', 'post'):
            if stage == 'aniso':
                if self.aniso > 0:
                    s = self._gate([B], self.aniso, torch.exp2(rng.randn(B, device=dev) * self.aniso_std), 1, dev)
                    if pct is not None:
                        s = torch.full_like(s, float(torch.exp2(torch.erfinv(pct * 2 - 1) * self.aniso_std)))
                    G = mul(_scale2(1 / s, s))
            elif self.rotate > 0:
                theta = (rng.rand(B, device=dev) * 2 - 1) * math.pi * self.rotate_max
                theta = torch.where(rng.rand(B, device=dev) < p_rot, theta, torch.zeros_like(theta))
                if pct is not None:
                    theta = torch.full_like(theta, float((pct * 2 - 1) * math.pi * self.rotate_max)) if stage == 'pre' else torch.zeros_like(theta)
                G = mul(_rotate2(theta))
        if self.xfrac > 0:
            t = self._gate([B, 1], self.xfrac, rng.randn(B, 2, device=dev) * self.xfrac_std, 0, dev)
            if pct is not None:
                t = torch.full_like(t, float(torch.erfinv(pct * 2 - 1) * self.xfrac_std))
            G = mul(_translate2(-t[:, 0] * width, -t[:, 1] * height))
        return G

    def _color(self, B, channels, dev, pct):
        C = None                                    # C @ color_in -> color_out (homogeneous 4x4)
        eye = torch.eye(4, device=dev)
        mul = lambda M: M if C is None else M @ C
        axis = torch.tensor([1, 1, 1, 0], dtype=torch.float32, device=dev) / math.sqrt(3)      # luma axis
        vv = torch.outer(axis, axis)
        if self.brightness > 0:
            b = self._gate([B], self.brightness, rng.randn(B, device=dev) * self.brightness_std, 0, dev)
            if pct is not None:
                b = torch.full_like(b, float(torch.erfinv(pct * 2 - 1) * self.brightness_std))
            C = mul(_mat([[1, 0, 0, b], [0, 1, 0, b], [0, 0, 1, b], [0, 0, 0, 1]], b))
        if self.contrast > 0:
            c = self._gate([B], self.contrast, torch.exp2(rng.randn(B, device=dev) * self.contrast_std), 1, dev)
            if pct is not None:
                c = torch.full_like(c, float(torch.exp2(torch.erfinv(pct * 2 - 1) * self.contrast_std)))
            C = mul(_mat([[c, 0, 0, 0], [0, c, 0, 0], [0, 0, c, 0], [0, 0, 0, 1]], c))
        if self.lumaflip > 0:
            i = self._gate([B, 1, 1], self.lumaflip, torch.floor(rng.rand(B, 1, 1, device=dev) * 2), 0, dev)
            if pct is not None:
                i = torch.full_like(i, float(torch.floor(pct * 2)))
            C = mul(eye - 2 * vv * i)               # Householder reflection about the luma axis
        if self.hue > 0 and channels > 1:
            theta = self._gate([B], self.hue, (rng.rand(B, device=dev) * 2 - 1) * math.pi * self.hue_max, 0, dev)
            if pct is not None:
                theta = torch.full_like(theta, float((pct * 2 - 1) * math.pi * self.hue_max))
            x, y, z = (float(v) for v in axis[:3])
            s, c = torch.sin(theta), torch.cos(theta)
            k = 1 - c
            C = mul(_mat([[x * x * k + c, x * y * k - z * s, x * z * k + y * s, 0],
                          [y * x * k + z * s, y * y * k + c, y * z * k - x * s, 0],
                          [z * x * k - y * s, z * y * k + x * s, z * z * k + c, 0],
                          [0, 0, 0, 1]], theta))     # rotation by theta about the luma axis
        if self.saturation > 0 and channels > 1:
            s = self._gate([B, 1, 1], self.saturation, torch.exp2(rng.randn(B, 1, 1, device=dev) * self.saturation_std), 1, dev)
            if pct is not None:
                s = torch.full_like(s, float(torch.exp2(torch.erfinv(pct * 2 - 1) * self.saturation_std)))
            C = mul(vv + (eye - vv) * s)
        return C

    # ---- execution ----------------------------------------------------------------------------------------------------
    def _warp(self, images, G):
        """Resample `images` through the inverse transforms G [B,3,3] (pixel units, origin at the image centre):
        reflect-pad by the margin the warped corners + the filter support need, up-sample x2, inverse-warp, down-sample x2."""
        B, ch, height, width = images.shape
        dev = images.device
        hz_pad = self.Hz_geom.shape[0] // 4
        cx, cy = (width - 1) / 2, (height - 1) / 2
        corners = torch.tensor([[-cx, -cy, 1], [cx, -cy, 1], [cx, cy, 1], [-cx, cy, 1]], dtype=torch.float32, device=dev)
        moved = G @ corners.t()                                              # [B, xyz, corner]
        ext = moved[:, :2, :].permute(1, 0, 2).flatten(1)                    # [xy, B * corner]
        margin = torch.cat([-ext, ext]).max(dim=1).values                    # [x0, y0, x1, y1]
        margin = margin + torch.tensor([hz_pad * 2 - cx, hz_pad * 2 - cy] * 2, dtype=torch.float32, device=dev)
        margin = margin.max(torch.zeros(4, device=dev)).min(torch.tensor([width - 1, height - 1] * 2, dtype=torch.float32, device=dev))
        mx0, my0, mx1, my1 = (int(v) for v in margin.ceil().to(torch.int32).tolist())       # one host read-back, as :281
        images = reflect_pad(images, [mx0, mx1, my0, my1])
        G = _translate2(torch.full([1], (mx0 - mx1) / 2, device=dev), torch.full([1], (my0 - my1) / 2, device=dev)) @ G
        images = upfirdn2d.upsample2d(x=images, f=self.Hz_geom, up=2)
        one = torch.ones([1], device=dev)
        G = _scale2(2 * one, 2 * one) @ G @ _scale2(one / 2, one / 2)
        G = _translate2(-0.5 * one, -0.5 * one) @ G @ _translate2(0.5 * one, 0.5 * one)
        out_shape = [B, ch, (height + hz_pad * 2) * 2, (width + hz_pad * 2) * 2]
        G = _scale2(one * (2 / images.shape[3]), one * (2 / images.shape[2])) @ G @ _scale2(one * (out_shape[3] / 2), one * (out_shape[2] / 2))
        images = affine_grid_sample(images, G[:, :2, :], out_shape)
        return upfirdn2d.downsample2d(x=images, f=self.Hz_geom, down=2, padding=-hz_pad * 2, flip_filter=True)

    def _band_filter(self, images, dev, pct):
        B, ch, height, width = images.shape
        bands = self.Hz_fbank.shape[0]
        assert len(self.imgfilter_bands) == bands
        power = torch.tensor([10, 1, 1, 1], dtype=torch.float32, device=dev) / 13         # expected 1/f power spectrum
        gains = torch.ones([B, bands], device=dev)
        for i, strength in enumerate(self.imgfilter_bands):
            t_i = self._gate([B], self.imgfilter * strength, torch.exp2(rng.randn(B, device=dev) * self.imgfilter_std), 1, dev)
            if pct is not None:
                t_i = torch.full_like(t_i, float(torch.exp2(torch.erfinv(pct * 2 - 1) * self.imgfilter_std))) if strength > 0 else torch.ones_like(t_i)
            t = torch.ones([B, bands], device=dev)
            t[:, i] = t_i
            gains = gains * (t / (power * t.square()).sum(dim=-1, keepdim=True).sqrt())
        taps = gains @ self.Hz_fbank                                         # [B, taps]: one separable filter per sample
        pad = self.Hz_fbank.shape[1] // 2
        images = reflect_pad(images, [pad, pad, pad, pad])
        return torch.cat([upfirdn2d.upfirdn2d(images[b:b + 1], taps[b], flip_filter=True) for b in range(B)], dim=0)

    def forward(self, images, debug_percentile=None):
        assert isinstance(images, torch.Tensor) and images.ndim == 4
        B, ch, height, width = images.shape
        dev = images.device
        pct = None if debug_percentile is None else torch.as_tensor(debug_percentile, dtype=torch.float32, device=dev)
        G = self._geometry(B, width, height, dev, pct)
        if G is not None:
            images = self._warp(images, G)
        C = self._color(B, ch, dev, pct)
        if C is not None:
            if ch == 3:
                images = color_affine(images, C.expand(B, 4, 4))
            elif ch == 1:
                row = C[:, :3, :].mean(dim=1, keepdim=True)
                images = images * row[:, :, :3].sum(dim=2, keepdim=True).unsqueeze(-1) + row[:, :, 3:].unsqueeze(-1)
            else:
                raise ValueError('Image must be RGB (3 channels) or L (1 channel)')
        if self.imgfilter > 0:
            images = self._band_filter(images, dev, pct)
        if self.noise > 0:
            sigma = self._gate([B, 1, 1, 1], self.noise, rng.randn(B, 1, 1, 1, device=dev).abs() * self.noise_std, 0, dev)
            if pct is not None:
                sigma = torch.full_like(sigma, float(torch.erfinv(pct) * self.noise_std))
            images = images + rng.randn(B, ch, height, width, device=dev) * sigma
        if self.cutout > 0:
            size = self._gate([B, 1, 1, 1, 1], self.cutout, torch.full([B, 2, 1, 1, 1], self.cutout_size, device=dev), 0, dev)
            center = rng.rand(B, 2, 1, 1, 1, device=dev)
            if pct is not None:
                size, center = torch.full_like(size, self.cutout_size), torch.full_like(center, float(pct))
            xs = (torch.arange(width, device=dev).reshape(1, 1, 1, -1) + 0.5) / width
            ys = (torch.arange(height, device=dev).reshape(1, 1, -1, 1) + 0.5) / height
            keep = torch.logical_or((xs - center[:, 0]).abs() >= size[:, 0] / 2, (ys - center[:, 1]).abs() >= size[:, 1] / 2)
            images = images * keep.to(torch.float32)
        return images


class ADA(AugmentPipe):
    """AugmentPipe whose strength p follows the discriminator's overfitting heuristic r_t = E[sign(D(real))]
    (nnutils/ada.py:5-36): every `interval` calls, p moves by batch * interval / (target_kimg * 1000) towards r_t > threshold."""

    def __init__(self, batch_size: int, interval: int = 4, target_kimg: int = 500, threshold: float = 0.6, **augment_kwargs) -> None:
        if augment_kwargs == {}:
            augment_kwargs = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1,
                                  brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1)
        super().__init__(**augment_kwargs)
        self._batch_size, self._interval, self._threshold = batch_size, interval, threshold
        self._target_img = target_kimg * 1000
        self._p_delta = batch_size * interval / self._target_img
        self._num_iter = 0
        self.register_buffer('signsum', torch.zeros([]))
        self.p.copy_(torch.zeros([]))

    @torch.no_grad()
    def update_p(self, prob: torch.Tensor):
        self.signsum.add_(torch.sign(prob).sum())
        self._num_iter += 1
        if self._num_iter == self._interval:
            r_t = self.signsum / (self._batch_size * self._interval)
            self.p.copy_((self.p + torch.sign(r_t - self._threshold) * self._p_delta).clamp_(0., 1.))
            self._num_iter = 0
            self.signsum.fill_(0.)
