"""DiffAugment ('color', 'translation', 'cutout') with the reference's signature and draw order.

Reference: thirdparty/diffaugment/DiffAugment.py:10-77.  Same arithmetic and the same sequence of random
draws (3 x rand(B,1,1,1) for colour, 2 x randint for translation / cutout); the translation is a pair of
1-D gathers on a zero-padded copy instead of the reference's NHWC permute + 3-D advanced indexing
(~10 passes over the batch), and stays differentiable w.r.t. x.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import rng


def DiffAugment(x, policy='', channels_first=True):
    if policy:
        if not channels_first:
            x = x.permute(0, 3, 1, 2)
        for p in policy.split(','):
            for f in AUGMENT_FNS[p]:
                x = f(x)
        if not channels_first:
            x = x.permute(0, 2, 3, 1)
        x = x.contiguous()
    return x


def rand_brightness(x):
    return x + (rng.rand(x.size(0), 1, 1, 1, device=x.device, dtype=x.dtype) - 0.5)


def rand_saturation(x):
    m = x.mean(dim=1, keepdim=True)
    return (x - m) * (rng.rand(x.size(0), 1, 1, 1, device=x.device, dtype=x.dtype) * 2) + m


def rand_contrast(x):
    m = x.mean(dim=[1, 2, 3], keepdim=True)
    return (x - m) * (rng.rand(x.size(0), 1, 1, 1, device=x.device, dtype=x.dtype) + 0.5) + m


def rand_translation(x, ratio=0.125):
    B, C, H, W = x.shape
    sh, sw = int(H * ratio + 0.5), int(W * ratio + 0.5)
    th = rng.randint(-sh, sh + 1, (B, 1, 1), device=x.device).reshape(B)
    tw = rng.randint(-sw, sw + 1, (B, 1, 1), device=x.device).reshape(B)
    rows = (torch.arange(H, device=x.device)[None, :] + th[:, None] + 1).clamp_(0, H + 1)     # [B,H] into padded
    cols = (torch.arange(W, device=x.device)[None, :] + tw[:, None] + 1).clamp_(0, W + 1)     # [B,W]
    xp = F.pad(x, [1, 1, 1, 1])
    xp = torch.gather(xp, 2, rows[:, None, :, None].expand(B, C, H, W + 2))
    return torch.gather(xp, 3, cols[:, None, None, :].expand(B, C, H, W))


def rand_cutout(x, ratio=0.5):
    B, C, H, W = x.shape
    ch, cw = int(H * ratio + 0.5), int(W * ratio + 0.5)
    oy = rng.randint(0, H + (1 - ch % 2), (B, 1, 1), device=x.device).reshape(B, 1)
    ox = rng.randint(0, W + (1 - cw % 2), (B, 1, 1), device=x.device).reshape(B, 1)
    ys = torch.arange(H, device=x.device)[None, :]
    xs = torch.arange(W, device=x.device)[None, :]
    # the reference clamps the cut window's indices into the image, so the window is clipped at the borders
    y0, y1 = (oy - ch // 2).clamp(0, H - 1), (oy - ch // 2 + ch - 1).clamp(0, H - 1)
    x0, x1 = (ox - cw // 2).clamp(0, W - 1), (ox - cw // 2 + cw - 1).clamp(0, W - 1)
    inside = ((ys >= y0) & (ys <= y1))[:, :, None] & ((xs >= x0) & (xs <= x1))[:, None, :]
    return x * (~inside).to(x.dtype).unsqueeze(1)


AUGMENT_FNS = {
    'color': [rand_brightness, rand_saturation, rand_contrast],
    'translation': [rand_translation],
    'cutout': [rand_cutout],
}
