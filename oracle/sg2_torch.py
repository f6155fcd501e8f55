"""CPU ORACLE (test infrastructure, NOT a product path) -- plain-PyTorch fp32 restatement of the reference's
StyleGAN2 generator, discriminator, DiffAugment and training-step body.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The reference path is pure PyTorch, so its restatement is written with stock torch ops; it is
functional (a flat ``state_dict`` in, tensors out) and does not import anything from animeface_b200.

Restated from (paths relative to the STomoya/animeface checkout):
  generator      implementations/StyleGAN2/model.py:71-135 (MapLinear, InjectNoise, ModulatedConv2d), :154-180,
                 :239-363 (ToImage, PixelNorm, Mapping, Synthesis, Generator)
  discriminator  implementations/StyleGAN2/model.py:186-236, :370-401
  step           implementations/StyleGAN2/utils.py:53-116 (loop body), :208-221 (Adam set-up),
                 :18-33 (path-length penalty and its running mean)
  losses         nnutils/loss/gan.py:98-114, nnutils/loss/penalty.py:11-26, 85-101
  augmentation   thirdparty/diffaugment/DiffAugment.py:10-53
  ema            nnutils/training.py:23-40

PARITY PIN: no reference test or fixture exists for this path; this file is pinned against the
reference itself by tests/golden/make_golden.py (run in the build container, imports /root/reference) and
checked by tests/test_oracle_golden.py.

Randomness is explicit: every function takes the random draws as arguments (``Draws``), in the order the
reference consumes them from the global generator, so CPU-reference, oracle and GPU runs can share them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

SLOPE = 0.2


# --------------------------------------------------------------------------------------------------------
# random draws

@dataclass
class Draws:
    """FIFO of pre-drawn random tensors, consumed in reference order."""
    items: list = field(default_factory=list)
    pos: int = 0

    def pop(self, shape=None):
        t = self.items[self.pos]
        self.pos += 1
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape), self.pos)
        return t


class FreshDraws:
    """Draws from the global torch generator (used by the CPU baseline timing)."""

    def __init__(self, device='cpu'):
        self.device = device

    def randn(self, *shape):
        return torch.randn(*shape, device=self.device)

    def rand(self, *shape):
        return torch.rand(*shape, device=self.device)

    def randint(self, lo, hi, shape):
        return torch.randint(lo, hi, shape, device=self.device)


class ReplayDraws:
    """Adapter giving a Draws FIFO the FreshDraws interface."""

    def __init__(self, draws: Draws):
        self.d = draws

    def randn(self, *shape):
        return self.d.pop(shape)

    def rand(self, *shape):
        return self.d.pop(shape)

    def randint(self, lo, hi, shape):
        return self.d.pop(shape)


# --------------------------------------------------------------------------------------------------------
# generator

def _elr_linear(sd, prefix, x):
    w, b = sd[prefix + '.layer.weight'], sd[prefix + '.layer.bias']
    return F.linear(x * (1.0 / math.sqrt(w.shape[1])), w, b)


def mapping(sd, z, map_lr=0.01, normalize=True):
    x = z
    if normalize:
        x = x / (x.pow(2).mean(dim=1, keepdim=True).sqrt() + 1e-4)
    i = 0
    while f'map.map.{i}.linear.layer.weight' in sd:
        x = F.leaky_relu(_elr_linear(sd, f'map.map.{i}.linear', x) * map_lr, SLOPE)
        i += 2
    return x


def modconv(sd, prefix, x, style, demod=True):
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    co, ci, k, _ = w.shape
    s = _elr_linear(sd, prefix + '.affine', style) + 1
    wm = w[None] * s[:, None, :, None, None] * (1.0 / math.sqrt(ci * k * k))
    if demod:
        wm = wm * torch.rsqrt(wm.pow(2).sum([2, 3, 4], keepdim=True) + 1e-4)
    B, _, H, W = x.shape
    y = F.conv2d(x.reshape(1, B * ci, H, W), wm.reshape(B * co, ci, k, k), padding=(k - 1) // 2, groups=B)
    return y.reshape(B, co, H, W) + b


def _up(x):
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)


def _blur(x):
    k = torch.tensor([1., 2., 1.], dtype=x.dtype, device=x.device)
    k = (k[:, None] * k[None, :] / 16).expand(x.shape[1], 1, 3, 3)
    return F.conv2d(x, k, padding=1, groups=x.shape[1])


def generator(sd, z, rng, map_lr=0.01, normalize=True):
    """Returns (image, style).  ``rng.randn(B,1,H,W)`` is called once per InjectNoise, in forward order."""
    style = mapping(sd, z, map_lr, normalize)
    B = z.shape[0]
    x = sd['const'].expand(B, -1, -1, -1)
    x = modconv(sd, 'synthesis.input', x, style)
    img = _up(modconv(sd, 'synthesis.input_to_image.conv', x, style, demod=False))
    nblocks = 0
    while f'synthesis.blocks.{nblocks}.block.2.weight' in sd:
        nblocks += 1
    for i in range(nblocks):
        p = f'synthesis.blocks.{i}.block'
        x = _blur(_up(x))
        j = 2
        while f'{p}.{j}.weight' in sd:
            x = modconv(sd, f'{p}.{j}', x, style)
            x = F.leaky_relu(x + rng.randn(B, 1, x.shape[2], x.shape[3]), SLOPE)
            j += 3
        img = modconv(sd, f'synthesis.to_images.{i}.conv', x, style, demod=False) + img
        if i + 1 < nblocks:
            img = _up(img)
    return torch.tanh(img), style


# --------------------------------------------------------------------------------------------------------
# discriminator

def _elr_conv(sd, prefix, x, pad):
    w, b = sd[prefix + '.layer.weight'], sd[prefix + '.layer.bias']
    return F.conv2d(x * (1.0 / math.sqrt(w[0].numel())), w, b, padding=pad)


def mbstd(x, group_size=4, eps=1e-4):
    B, C, H, W = x.shape
    G = group_size if B % group_size == 0 else B
    y = x.reshape(G, -1, C, H, W)
    y = y - y.mean(0, keepdim=True)
    y = (y.square().mean(0) + eps).sqrt().mean([1, 2, 3], keepdim=True)
    return torch.cat([x, y.repeat(G, 1, H, W)], dim=1)


def discriminator(sd, x, mbsd_groups=4):
    x = F.leaky_relu(_elr_conv(sd, 'from_rgb.0', x, 0), SLOPE)
    i = 0
    while f'blocks.{i}.skip.layer.weight' in sd:
        p = f'blocks.{i}'
        t = _elr_conv(sd, p + '.skip', x, 0)
        j = 0
        while f'{p}.block.{j}.layer.weight' in sd:
            x = F.leaky_relu(_elr_conv(sd, f'{p}.block.{j}', x, 1), SLOPE)
            j += 2
        x = (F.avg_pool2d(x, 2) + F.avg_pool2d(t, 2)) / math.sqrt(2)
        i += 1
    x = mbstd(x, mbsd_groups)
    x = F.leaky_relu(_elr_conv(sd, f'blocks.{i + 1}', x, 1), SLOPE)
    x = x.reshape(x.shape[0], -1)
    x = F.leaky_relu(_elr_linear(sd, f'blocks.{i + 4}', x), SLOPE)
    return _elr_linear(sd, f'blocks.{i + 6}', x)


# --------------------------------------------------------------------------------------------------------
# DiffAugment 'color,translation'

def diffaugment(x, rng, policy='color,translation'):
    """Per-sample brightness, saturation, contrast, then an integer shift of up to 1/8 of the size with zero
    fill.  Draw order: rand(B,1,1,1) x3, randint(-s, s+1, (B,1,1)) x2 (x shift along H first)."""
    B, C, H, W = x.shape
    for p in policy.split(',') if policy else []:
        if p == 'color':
            x = x + (rng.rand(B, 1, 1, 1) - 0.5)
            m = x.mean(dim=1, keepdim=True)
            x = (x - m) * (rng.rand(B, 1, 1, 1) * 2) + m
            m = x.mean(dim=[1, 2, 3], keepdim=True)
            x = (x - m) * (rng.rand(B, 1, 1, 1) + 0.5) + m
        elif p == 'translation':
            sh, sw = int(H * 0.125 + 0.5), int(W * 0.125 + 0.5)
            th = rng.randint(-sh, sh + 1, (B, 1, 1)).reshape(B)
            tw = rng.randint(-sw, sw + 1, (B, 1, 1)).reshape(B)
            # output[y, x] = input[y + th, x + tw] (zero outside), expressed as a gather on a zero-padded copy
            ys = (torch.arange(H, device=x.device)[None, :] + th[:, None] + 1).clamp(0, H + 1)      # [B,H]
            xs = (torch.arange(W, device=x.device)[None, :] + tw[:, None] + 1).clamp(0, W + 1)      # [B,W]
            xp = F.pad(x, [1, 1, 1, 1])
            xp = torch.gather(xp, 2, ys[:, None, :, None].expand(B, C, H, W + 2))
            x = torch.gather(xp, 3, xs[:, None, None, :].expand(B, C, H, W))
        else:
            raise NotImplementedError(p)
    return x.contiguous()


# --------------------------------------------------------------------------------------------------------
# losses / penalties

def d_loss_ns(real_logits, fake_logits):
    return F.softplus(-real_logits).mean() + F.softplus(fake_logits).mean()


def g_loss_ns(fake_logits):
    return F.softplus(-fake_logits).mean()


def r1_penalty(sd_d, real, mbsd_groups=4):
    x = real.detach().requires_grad_(True)
    out = discriminator(sd_d, x, mbsd_groups)
    g, = torch.autograd.grad(out, x, torch.ones_like(out), create_graph=True, retain_graph=True)
    return g.reshape(g.shape[0], -1).norm(2, dim=1).pow(2).mean() / 2.


def pl_penalty(style, image, pl_mean, rng):
    """implementations/StyleGAN2/utils.py:18-29: noise = randn(image.shape)/sqrt(H*W); g = d sum(image*noise)/d style
    (create_graph); mean_b((||g_b||_2 - pl_mean)^2)."""
    noise = rng.randn(*image.shape) / math.sqrt(image.shape[2] * image.shape[3])
    g, = torch.autograd.grad((image * noise).sum(), style, create_graph=True, retain_graph=True)
    return (g.pow(2).sum(dim=1).sqrt() - pl_mean).pow(2).mean()


# --------------------------------------------------------------------------------------------------------
# one training step (loop body of implementations/StyleGAN2/utils.py:53-116, AMP off)

@dataclass
class StepConfig:
    latent_dim: int = 512
    r1_lambda: float = 10.
    d_k: int = 16
    policy: str = 'color,translation'
    mbsd_groups: int = 4
    map_lr: float = 0.01
    lr: float = 1e-3
    betas: tuple = (0., 0.99)
    ema_decay: float = 0.999
    pl_lambda: float = 0.
    g_k: int = 8


def adam_hparams(cfg: StepConfig):
    """(g_lr, g_betas, d_lr, d_betas) as utils.py:208-218 (lazy-regularisation ratios k/(k+1) when the penalty is on)."""
    def scaled(on, k):
        r = k / (k + 1) if on else 1.0
        return cfg.lr * r, (cfg.betas[0] ** r, cfg.betas[1] ** r)
    g_lr, g_b = scaled(cfg.pl_lambda > 0, cfg.g_k)
    d_lr, d_b = scaled(cfg.r1_lambda > 0, cfg.d_k)
    return g_lr, g_b, d_lr, d_b


def train_step(sd_g, sd_d, sd_ema, opt_g, opt_d, real, step_idx, rng, cfg: StepConfig, state=None):
    """sd_* are dicts of leaf tensors (requires_grad=True for G and D); opt_* are torch optimizers over their
    values.  ``state`` (a dict) carries ``pl_mean`` across steps when cfg.pl_lambda > 0.
    Returns (D_loss, G_loss, fake) as detached tensors."""
    B = real.shape[0]
    opt_g.zero_grad()
    opt_d.zero_grad()
    # --- discriminator phase
    z = rng.randn(B, cfg.latent_dim)
    real_prob = discriminator(sd_d, diffaugment(real, rng, cfg.policy), cfg.mbsd_groups)
    fake, _ = generator(sd_g, z, rng, cfg.map_lr)
    fake_prob = discriminator(sd_d, diffaugment(fake, rng, cfg.policy).detach(), cfg.mbsd_groups)
    if step_idx % cfg.d_k == 0 and cfg.r1_lambda > 0 and step_idx != 0:
        d_loss = r1_penalty(sd_d, real, cfg.mbsd_groups) * cfg.r1_lambda * cfg.d_k
    else:
        d_loss = d_loss_ns(real_prob, fake_prob)
    d_loss.backward()
    opt_d.step()
    # --- generator phase
    z = rng.randn(B, cfg.latent_dim)
    fake, style = generator(sd_g, z, rng, cfg.map_lr)
    fake_prob = discriminator(sd_d, diffaugment(fake, rng, cfg.policy), cfg.mbsd_groups)
    if step_idx % cfg.g_k == 0 and cfg.pl_lambda > 0 and step_idx != 0:
        pl = pl_penalty(style, fake, state['pl_mean'], rng)
        g_loss = pl * cfg.pl_lambda * cfg.g_k
        state['pl_mean'] = 0.99 * state['pl_mean'] + 0.01 * float(pl.detach())      # EMA of the penalty, utils.py:100-103
    else:
        g_loss = g_loss_ns(fake_prob)
    g_loss.backward()
    opt_g.step()
    if sd_ema is not None:
        with torch.no_grad():
            for k, v in sd_ema.items():
                if k in sd_g and sd_g[k].requires_grad:
                    v.mul_(cfg.ema_decay).add_(sd_g[k].detach(), alpha=1 - cfg.ema_decay)
    return d_loss.detach(), g_loss.detach(), fake.detach()


# --------------------------------------------------------------------------------------------------------
# random-init state dicts with the reference's shapes and init (utils.py:186-201; model.py:335-349, 404-408)

def init_generator_sd(image_size=256, image_channels=3, style_dim=512, channels=32, max_channels=512,
                      num_conv=2, map_layers=8, map_lr=0.01, gen=None):
    def n(*shape, std=1.0):
        return torch.randn(*shape, generator=gen) * std
    sd = {}
    for i in range(map_layers):
        sd[f'map.map.{2 * i}.linear.layer.weight'] = n(style_dim, style_dim, std=1 / map_lr)
        sd[f'map.map.{2 * i}.linear.layer.bias'] = torch.zeros(style_dim)

    def mod(prefix, ci, co, k):
        sd[prefix + '.affine.layer.weight'] = n(ci, style_dim)
        sd[prefix + '.affine.layer.bias'] = torch.zeros(ci)
        sd[prefix + '.weight'] = n(co, ci, k, k)
        sd[prefix + '.bias'] = torch.zeros(1, co, 1, 1)
    ch = channels * (2 ** int(math.log2(image_size) - 2))
    och = min(max_channels, ch)
    mod('synthesis.input', style_dim, och, 3)
    mod('synthesis.input_to_image.conv', och, image_channels, 1)
    resl, i = 4, 0
    while resl < image_size:
        resl *= 2
        ch //= 2
        ich, och = och, min(max_channels, ch)
        sd[f'synthesis.blocks.{i}.block.1.kernel'] = torch.tensor([[[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]]) / 16
        for c in range(num_conv):
            mod(f'synthesis.blocks.{i}.block.{2 + 3 * c}', ich if c == 0 else och, och, 3)
            sd[f'synthesis.blocks.{i}.block.{3 + 3 * c}.scale'] = torch.zeros(1)
        mod(f'synthesis.to_images.{i}.conv', och, image_channels, 1)
        i += 1
    sd['const'] = n(1, style_dim, 4, 4)
    return sd


def init_discriminator_sd(image_size=256, image_channels=3, channels=32, max_channels=512, num_conv=2, gen=None):
    def n(*shape):
        return torch.randn(*shape, generator=gen)
    sd = {}

    def conv(prefix, ci, co, k):
        sd[prefix + '.layer.weight'] = n(co, ci, k, k)
        sd[prefix + '.layer.bias'] = torch.zeros(co)
    och = channels
    conv('from_rgb.0', image_channels, och, 1)
    resl, i = image_size, 0
    while resl > 4:
        resl //= 2
        channels *= 2
        ich, och = och, min(max_channels, channels)
        for c in range(num_conv):
            conv(f'blocks.{i}.block.{2 * c}', ich if c == 0 else och, och, 3)
        conv(f'blocks.{i}.skip', ich, och, 1)
        i += 1
    conv(f'blocks.{i + 1}', och + 1, och, 3)
    sd[f'blocks.{i + 4}.layer.weight'] = n(och, och * resl * resl)
    sd[f'blocks.{i + 4}.layer.bias'] = torch.zeros(och)
    sd[f'blocks.{i + 6}.layer.weight'] = n(1, och)
    sd[f'blocks.{i + 6}.layer.bias'] = torch.zeros(1)
    return sd
