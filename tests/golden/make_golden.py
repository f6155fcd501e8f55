"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE (STomoya/animeface) on CPU.

Run in the build container only (it imports /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or fixtures for this path, so these files are the parity pin:
outputs of the reference's own code (torch CPU fp32) on seeded inputs, with every random draw recorded so
that the oracle and the GPU path can replay them.  Files (all small, committed):
    ops.npz      upfirdn2d / bias_act reference ops (thirdparty/stylegan3_ops, impl='ref') incl. gradients
    modules.npz  single StyleGAN2 modules (implementations/StyleGAN2/model.py) forward + gradients
    model.npz    small G/D: images, logits, losses, all parameter gradients, R1, 3-step trajectory
    sg3d.npz     StyleGAN3-style discriminator (implementations/StyleGAN3/model.py:382-510): logits, D-loss and R1
                 parameter gradients, conv2d_resample cases
    sg3g.npz     StyleGAN3 generator (implementations/StyleGAN3/model.py:32-380) and filtered_lrelu cases incl. gradients
    ada.npz      AugmentPipe / ADA (thirdparty/ada/augment.py, nnutils/ada.py): outputs, gradients, recorded draws, p updates
    resample.npz conv2d_resample up-sampling / grouped branches and conv_transpose2d incl. gradients
    pl.npz       path-length penalty (implementations/StyleGAN2/utils.py:18-33): value, per-sample gradient norms,
                 all second-order parameter gradients, 4-step trajectory with a PL step and an R1 step
"""
import functools
import os
import sys

import numpy as np
import torch

REF = '/root/reference'
sys.path.insert(0, REF)
os.chdir(REF)
HERE = os.path.dirname(os.path.abspath(__file__)) if '__file__' in globals() else '/root/repo/tests/golden'

from implementations.StyleGAN2 import model as ref_model                      # noqa: E402
from nnutils.loss import NonSaturatingLoss, r1_regularizer                     # noqa: E402
from nnutils.training import update_ema                                        # noqa: E402
from thirdparty.diffaugment import DiffAugment                                 # noqa: E402
from thirdparty.stylegan3_ops.ops import bias_act as ref_ba                    # noqa: E402
from thirdparty.stylegan3_ops.ops import upfirdn2d as ref_up                   # noqa: E402

torch.set_num_threads(8)
A = lambda t: t.detach().cpu().numpy().copy()


# ------------------------------------------------------------------------------------------------ ops
def gen_ops():
    out = {}
    g = torch.Generator().manual_seed(1234)
    # (name, x shape, filter spec, kwargs, wrapper)
    cases = [
        ('up2_f4', (2, 3, 9, 7), [1, 3, 3, 1], dict(up=2), 'upsample2d'),
        ('down2_f4', (2, 4, 12, 10), [1, 3, 3, 1], dict(down=2), 'downsample2d'),
        ('filt_f4', (1, 4, 8, 8), [1, 3, 3, 1], dict(), 'filter2d'),
        ('filt_f3_flip', (1, 2, 6, 9), [1, 2, 5], dict(flip_filter=True, gain=1.5), 'filter2d'),
        ('raw_pad_neg', (1, 2, 10, 10), [1, 3, 3, 1], dict(up=1, down=1, padding=[-1, 2, 0, -2]), 'upfirdn2d'),
        ('raw_up3_down2', (1, 2, 5, 6), [1, 4, 6, 4, 1], dict(up=3, down=2, padding=[2, 1, 3, 0], gain=2.0), 'upfirdn2d'),
        ('raw_xy', (1, 3, 6, 5), [1, 3, 3, 1], dict(up=[2, 1], down=[1, 2], padding=[1, 1, 2, 1]), 'upfirdn2d'),
        ('sep12_up2', (1, 3, 10, 10), list(np.hanning(14)[1:-1]), dict(up=2), 'upsample2d'),
        ('sep12_down2', (1, 3, 20, 20), list(np.hanning(14)[1:-1]), dict(down=2, padding=-2, flip_filter=True), 'downsample2d'),
        ('avg_as_down2', (2, 4, 8, 8), [1, 1], dict(down=2), 'downsample2d'),
        ('none_filter', (1, 2, 4, 4), None, dict(up=2), 'upfirdn2d'),
    ]
    names = []
    for name, shape, fspec, kw, wrap in cases:
        x = torch.randn(shape, generator=g).requires_grad_(True)
        f = None if fspec is None else ref_up.setup_filter(fspec)
        y = getattr(ref_up, wrap)(x, f, impl='ref', **kw)
        gy = torch.randn(y.shape, generator=g)
        gx, = torch.autograd.grad(y, x, gy)
        out[f'up.{name}.x'] = A(x); out[f'up.{name}.y'] = A(y); out[f'up.{name}.gy'] = A(gy); out[f'up.{name}.gx'] = A(gx)
        if f is not None:
            out[f'up.{name}.f'] = A(f)
        names.append(name)
    out['up.names'] = np.array(names)
    out['up.cases'] = np.array([repr((n, s, (None if f is None else [float(v) for v in f]), k, w)) for n, s, f, k, w in cases])
    # setup_filter variants
    for i, (spec, kw) in enumerate([([1, 3, 3, 1], {}), ([1, 2, 1], dict(gain=4)), (list(range(1, 10)), {}),
                                    ([1, 3, 3, 1], dict(flip_filter=True, normalize=False)), ([1, 2, 3], dict(separable=True, gain=2))]):
        out[f'sf.{i}'] = A(ref_up.setup_filter(spec, **kw))
    # bias_act: every activation, with bias (dim 1) and clamp; 1st and 2nd order gradients from autograd
    acts = list(ref_ba.activation_funcs.keys())
    out['ba.acts'] = np.array(acts)
    for act in acts:
        for variant, kw in (('plain', dict()), ('clamp', dict(clamp=0.7, gain=1.3, alpha=0.3))):
            x = (torch.randn((3, 4, 5, 2), generator=g) * 2).requires_grad_(True)
            b = torch.randn((4,), generator=g).requires_grad_(True)
            y = ref_ba.bias_act(x, b, dim=1, act=act, impl='ref', **kw)
            gy = torch.randn(y.shape, generator=g).requires_grad_(True)
            gx, gb = torch.autograd.grad(y, (x, b), gy, create_graph=True)
            ggx = torch.randn(gx.shape, generator=g)
            d_gy, d_x = torch.autograd.grad(gx, (gy, x), ggx, allow_unused=True)
            k = f'ba.{act}.{variant}'
            out[k + '.x'] = A(x); out[k + '.b'] = A(b); out[k + '.y'] = A(y); out[k + '.gy'] = A(gy)
            out[k + '.gx'] = A(gx); out[k + '.gb'] = A(gb); out[k + '.ggx'] = A(ggx); out[k + '.d_gy'] = A(d_gy)
            out[k + '.d_x'] = A(d_x) if d_x is not None else np.zeros(x.shape, np.float32)
    np.savez_compressed(os.path.join(HERE, 'ops.npz'), **out)
    print('ops.npz', len(out))


# -------------------------------------------------------------------------------------------- modules
def gen_modules():
    out = {}
    g = torch.Generator().manual_seed(4321)
    rn = lambda *s: torch.randn(*s, generator=g)
    # Upsample2x('bilinear') [+ Blur2d]: forward and adjoint, odd sizes and the 1-pixel edge case
    up, blur = ref_model.Upsample2x('bilinear'), ref_model.Blur2d()
    for name, shape in (('a', (2, 4, 5, 7)), ('b', (1, 8, 1, 3)), ('c', (1, 3, 4, 4))):
        x = rn(*shape).requires_grad_(True)
        for tag, fn in (('upblur', lambda t: blur(up(t))), ('up', up)):
            y = fn(x)
            gy = rn(*y.shape)
            gx, = torch.autograd.grad(y, x, gy)
            out[f'{tag}.{name}.x'] = A(x); out[f'{tag}.{name}.y'] = A(y); out[f'{tag}.{name}.gy'] = A(gy); out[f'{tag}.{name}.gx'] = A(gx)
    x = rn(2, 4, 6, 8)
    out['blur.x'] = A(x); out['blur.y'] = A(blur(x))
    out['avg.x'] = A(x); out['avg.y'] = A(ref_model.Downsample2x('avg')(x))
    # MiniBatchStdDev: forward, gradient, and gradient of (gradient . v) -- the R1 pattern
    for name, shape, gs in (('g4', (8, 6, 4, 4), 4), ('odd', (6, 4, 2, 2), 4)):
        m = ref_model.MiniBatchStdDev(gs)
        x = rn(*shape).requires_grad_(True)
        y = m(x)
        gy = rn(*y.shape).requires_grad_(True)
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        v = rn(*gx.shape)
        d_x, d_gy = torch.autograd.grad(gx, (x, gy), v)
        for k_, t in (('x', x), ('y', y), ('gy', gy), ('gx', gx), ('v', v), ('d_x', d_x), ('d_gy', d_gy)):
            out[f'mbstd.{name}.{k_}'] = A(t)
        out[f'mbstd.{name}.group'] = np.array(gs)
    # ModulatedConv2d (k3 demod, k1 no demod) forward + all gradients
    for name, (ci, co, sdim, k, demod, hw) in (('k3', (8, 12, 16, 3, True, 6)), ('k1', (8, 3, 16, 1, False, 5)),
                                                ('k3b', (16, 8, 16, 3, True, 4))):
        m = ref_model.ModulatedConv2d(ci, co, sdim, k, demod=demod)
        m.weight.data.copy_(rn(*m.weight.shape)); m.bias.data.copy_(rn(*m.bias.shape) * 0.3)
        m.affine.layer.weight.data.copy_(rn(*m.affine.layer.weight.shape)); m.affine.layer.bias.data.copy_(rn(ci) * 0.2)
        x = rn(3, ci, hw, hw).requires_grad_(True)
        w = rn(3, sdim).requires_grad_(True)
        y = m(x, w)
        gy = rn(*y.shape)
        grads = torch.autograd.grad(y, (x, w, m.weight, m.bias, m.affine.layer.weight, m.affine.layer.bias), gy)
        for k_, t in zip(('x', 'style', 'y', 'gy', 'weight', 'bias', 'aw', 'ab'),
                         (x, w, y, gy, m.weight, m.bias, m.affine.layer.weight, m.affine.layer.bias)):
            out[f'mod.{name}.{k_}'] = A(t)
        for k_, t in zip(('gx', 'gstyle', 'gweight', 'gbias', 'gaw', 'gab'), grads):
            out[f'mod.{name}.{k_}'] = A(t)
        out[f'mod.{name}.demod'] = np.array(demod)
    # DBlock forward + gradients + double backward (R1 pattern)
    # every parameter is drawn from the seeded generator `g` (init_weight_N01's N(0,1) weights; non-zero biases so their
    # gradients are exercised): the file regenerates bit for bit
    blk = ref_model.DBlock(8, 16, 2)
    for p in blk.parameters():
        p.data.copy_(rn(*p.shape) * (0.3 if p.ndim == 1 else 1.0))
    x = rn(4, 8, 8, 8).requires_grad_(True)
    y = blk(x)
    gy = rn(*y.shape)
    params = list(blk.parameters())
    grads = torch.autograd.grad(y, [x] + params, gy, create_graph=True)
    pen = grads[0].pow(2).sum()
    g2 = torch.autograd.grad(pen, params, allow_unused=True)
    out['dblock.x'] = A(x); out['dblock.y'] = A(y); out['dblock.gy'] = A(gy); out['dblock.gx'] = A(grads[0])
    for (n_, p), g1, gg in zip(blk.named_parameters(), grads[1:], g2):
        out[f'dblock.p.{n_}'] = A(p); out[f'dblock.g.{n_}'] = A(g1)
        out[f'dblock.g2.{n_}'] = A(gg) if gg is not None else np.zeros(p.shape, np.float32)
    out['dblock.names'] = np.array([n_ for n_, _ in blk.named_parameters()])
    np.savez_compressed(os.path.join(HERE, 'modules.npz'), **out)
    print('modules.npz', len(out))


# ---------------------------------------------------------------------------------------------- model
class Recorder:
    """Records every tensor drawn through torch.randn / rand / randint while active."""

    def __init__(self):
        self.items = []

    def __enter__(self):
        self._orig = (torch.randn, torch.rand, torch.randint)
        rec = self.items

        def wrap(fn):
            def inner(*a, **k):
                t = fn(*a, **k)
                rec.append(t.clone())
                return t
            return inner
        torch.randn, torch.rand, torch.randint = wrap(torch.randn), wrap(torch.rand), wrap(torch.randint)
        return self

    def __exit__(self, *exc):
        torch.randn, torch.rand, torch.randint = self._orig


CFG = dict(image_size=32, image_channels=3, style_dim=32, channels=8, max_channels=32, block_num_conv=2,
           map_num_layers=4, mbsd_groups=4, batch=8, d_k=2, r1_lambda=10., lr=1e-3, betas=(0., 0.99))


def gen_model():
    out = {'cfg': np.array(repr(CFG))}
    torch.manual_seed(777)
    c = CFG
    G = ref_model.Generator(c['image_size'], c['image_channels'], c['style_dim'], c['channels'], c['max_channels'],
                            c['block_num_conv'], c['map_num_layers'], True, 0.01)
    G_ema = ref_model.Generator(c['image_size'], c['image_channels'], c['style_dim'], c['channels'], c['max_channels'],
                                c['block_num_conv'], c['map_num_layers'], True, 0.01)
    D = ref_model.Discriminator(c['image_size'], c['image_channels'], c['channels'], c['max_channels'],
                                c['block_num_conv'], c['mbsd_groups'])
    G.init_weight(map_init_func=functools.partial(ref_model.init_weight_N01, lr=0.01), syn_init_func=ref_model.init_weight_N01)
    G_ema.eval()
    # the reference initialises G_ema with update_ema(decay=0) on torch.empty() storage (utils.py:199-200),
    # which yields NaN when that memory happens to hold NaN/inf; start from finite values so it is exact.
    G_ema.load_state_dict(G.state_dict())
    update_ema(G, G_ema, decay=0)
    D.apply(ref_model.init_weight_N01)
    for k, v in G.state_dict().items():
        out['G0.' + k] = A(v)
    for k, v in D.state_dict().items():
        out['D0.' + k] = A(v)
    B = c['batch']
    # --- single forward / backward with recorded draws
    z = torch.randn(B, c['style_dim'])
    real = torch.rand(B, 3, c['image_size'], c['image_size']) * 2 - 1
    out['z'] = A(z); out['real'] = A(real)
    with Recorder() as rec:
        image, style = G(z)
    for i, t in enumerate(rec.items):
        out[f'fwd.noise.{i}'] = A(t)
    out['fwd.n_noise'] = np.array(len(rec.items))
    out['fwd.image'] = A(image); out['fwd.style'] = A(style)
    logits_fake = D(image)
    logits_real = D(real)
    out['fwd.logits_fake'] = A(logits_fake); out['fwd.logits_real'] = A(logits_real)
    loss = NonSaturatingLoss()
    # G loss gradients (through D into G) and D loss gradients
    g_loss = loss.g_loss(logits_fake)
    gg = torch.autograd.grad(g_loss, [p for p in G.parameters()], retain_graph=True, allow_unused=True)
    for (n_, p), g_ in zip(G.named_parameters(), gg):
        out['ggrad.' + n_] = A(g_) if g_ is not None else np.zeros(p.shape, np.float32)
    out['g_loss'] = A(g_loss)
    d_loss = loss.d_loss(logits_real, D(image.detach()))
    dg = torch.autograd.grad(d_loss, list(D.parameters()), allow_unused=True)
    for (n_, p), g_ in zip(D.named_parameters(), dg):
        out['dgrad.' + n_] = A(g_) if g_ is not None else np.zeros(p.shape, np.float32)
    out['d_loss'] = A(d_loss)
    # R1 on the real batch: value, gradient w.r.t. input, parameter gradients (second order)
    r1 = r1_regularizer()(real, D, None)
    r1g = torch.autograd.grad(r1, list(D.parameters()), allow_unused=True)
    out['r1'] = A(r1)
    for (n_, p), g_ in zip(D.named_parameters(), r1g):
        out['r1grad.' + n_] = A(g_) if g_ is not None else np.zeros(p.shape, np.float32)
        out['r1none.' + n_] = np.array(g_ is None)
    # --- 3-step trajectory (loop body of implementations/StyleGAN2/utils.py:53-116, AMP off), d_k = 2 so
    # step 2 is an R1 step.  All draws recorded.
    ratio = c['d_k'] / (c['d_k'] + 1)
    opt_g = torch.optim.Adam(G.parameters(), lr=c['lr'], betas=c['betas'])
    opt_d = torch.optim.Adam(D.parameters(), lr=c['lr'] * ratio, betas=(c['betas'][0] ** ratio, c['betas'][1] ** ratio))
    augment = functools.partial(DiffAugment, policy='color,translation')
    r1_loss = r1_regularizer()
    steps = 3
    for it in range(steps):
        real = torch.rand(B, 3, c['image_size'], c['image_size']) * 2 - 1
        out[f'traj.{it}.real'] = A(real)
        with Recorder() as rec:
            opt_g.zero_grad(); opt_d.zero_grad()
            z = torch.randn(B, c['style_dim'])
            real_prob = D(augment(real))
            fake, _ = G(z)
            fake_prob = D(augment(fake).detach())
            if it % c['d_k'] == 0 and it != 0:
                D_loss = r1_loss(real, D, None) * c['r1_lambda'] * c['d_k']
            else:
                D_loss = loss.d_loss(real_prob, fake_prob)
            D_loss.backward(); opt_d.step()
            z = torch.randn(B, c['style_dim'])
            fake, style = G(z)
            fake_prob = D(augment(fake))
            G_loss = loss.g_loss(fake_prob)
            G_loss.backward(); opt_g.step()
            update_ema(G, G_ema)
        for i, t in enumerate(rec.items):
            out[f'traj.{it}.draw.{i}'] = A(t)
        out[f'traj.{it}.n_draws'] = np.array(len(rec.items))
        out[f'traj.{it}.d_loss'] = A(D_loss); out[f'traj.{it}.g_loss'] = A(G_loss); out[f'traj.{it}.fake'] = A(fake)
    for k, v in G.state_dict().items():
        out['G3.' + k] = A(v)
    for k, v in D.state_dict().items():
        out['D3.' + k] = A(v)
    for k, v in G_ema.state_dict().items():
        out['E3.' + k] = A(v)
    out['traj.steps'] = np.array(steps)
    np.savez_compressed(os.path.join(HERE, 'model.npz'), **out)
    print('model.npz', len(out))


def gen_pl():
    """Path-length regulariser through the reference's own pl_penalty / train-loop arithmetic."""
    from implementations.StyleGAN2.utils import pl_penalty, update_pl_mean
    c = dict(CFG, d_k=3, g_k=2, pl_lambda=2.)
    out = {'cfg': np.array(repr(c))}
    torch.manual_seed(4242)
    mk = lambda: ref_model.Generator(c['image_size'], c['image_channels'], c['style_dim'], c['channels'], c['max_channels'],
                                     c['block_num_conv'], c['map_num_layers'], True, 0.01)
    G, G_ema = mk(), mk()
    D = ref_model.Discriminator(c['image_size'], c['image_channels'], c['channels'], c['max_channels'],
                                c['block_num_conv'], c['mbsd_groups'])
    G.init_weight(map_init_func=functools.partial(ref_model.init_weight_N01, lr=0.01), syn_init_func=ref_model.init_weight_N01)
    G_ema.eval()
    G_ema.load_state_dict(G.state_dict())
    update_ema(G, G_ema, decay=0)
    D.apply(ref_model.init_weight_N01)
    for k, v in G.state_dict().items():
        out['G0.' + k] = A(v)
    for k, v in D.state_dict().items():
        out['D0.' + k] = A(v)
    B = c['batch']
    # --- one evaluation: value, per-sample norms, parameter gradients of the penalty (second order through G)
    z = torch.randn(B, c['style_dim'])
    out['z'] = A(z)
    pl_mean0 = 0.37
    out['pl_mean0'] = np.array(pl_mean0, np.float32)
    with Recorder() as rec:
        image, style = G(z)
        pl = pl_penalty(style, image, pl_mean0, None)
    for i, t in enumerate(rec.items):
        out[f'eval.draw.{i}'] = A(t)
    out['eval.n_draws'] = np.array(len(rec.items))
    out['eval.image'] = A(image); out['eval.pl'] = A(pl)
    noise = rec.items[-1] / np.sqrt(image.shape[2] * image.shape[3])
    gnorm = torch.autograd.grad((image * noise).sum(), style, retain_graph=True)[0].pow(2).sum(1).sqrt()
    out['eval.gnorm'] = A(gnorm)
    pg = torch.autograd.grad(pl, list(G.parameters()), allow_unused=True)
    for (n_, p), g_ in zip(G.named_parameters(), pg):
        out['plgrad.' + n_] = A(g_) if g_ is not None else np.zeros(p.shape, np.float32)
        out['plnone.' + n_] = np.array(g_ is None)
    # --- 4-step trajectory, loop body of utils.py:53-116 with pl_lambda > 0: step 2 is a PL step, step 3 an R1 step
    loss = NonSaturatingLoss()
    rg, rd = c['g_k'] / (c['g_k'] + 1), c['d_k'] / (c['d_k'] + 1)
    opt_g = torch.optim.Adam(G.parameters(), lr=c['lr'] * rg, betas=(c['betas'][0] ** rg, c['betas'][1] ** rg))
    opt_d = torch.optim.Adam(D.parameters(), lr=c['lr'] * rd, betas=(c['betas'][0] ** rd, c['betas'][1] ** rd))
    augment = functools.partial(DiffAugment, policy='color,translation')
    r1_loss = r1_regularizer()
    pl_mean = 0.
    steps = 4
    for it in range(steps):
        real = torch.rand(B, 3, c['image_size'], c['image_size']) * 2 - 1
        out[f'traj.{it}.real'] = A(real)
        with Recorder() as rec:
            opt_g.zero_grad(); opt_d.zero_grad()
            z = torch.randn(B, c['style_dim'])
            real_prob = D(augment(real))
            fake, _ = G(z)
            fake_prob = D(augment(fake).detach())
            if it % c['d_k'] == 0 and it != 0:
                D_loss = r1_loss(real, D, None) * c['r1_lambda'] * c['d_k']
            else:
                D_loss = loss.d_loss(real_prob, fake_prob)
            D_loss.backward(); opt_d.step()
            z = torch.randn(B, c['style_dim'])
            fake, style = G(z)
            fake_prob = D(augment(fake))
            if it % c['g_k'] == 0 and it != 0:
                pl = pl_penalty(style, fake, pl_mean, None)
                G_loss = pl * c['pl_lambda'] * c['g_k']
                pl_mean = update_pl_mean(pl_mean, np.mean(pl.detach().cpu().numpy()))
            else:
                G_loss = loss.g_loss(fake_prob)
            G_loss.backward(); opt_g.step()
            update_ema(G, G_ema)
        for i, t in enumerate(rec.items):
            out[f'traj.{it}.draw.{i}'] = A(t)
        out[f'traj.{it}.n_draws'] = np.array(len(rec.items))
        out[f'traj.{it}.d_loss'] = A(D_loss); out[f'traj.{it}.g_loss'] = A(G_loss); out[f'traj.{it}.fake'] = A(fake)
        out[f'traj.{it}.pl_mean'] = np.array(pl_mean, np.float32)
    for k, v in G.state_dict().items():
        out['G4.' + k] = A(v)
    for k, v in D.state_dict().items():
        out['D4.' + k] = A(v)
    for k, v in G_ema.state_dict().items():
        out['E4.' + k] = A(v)
    out['traj.steps'] = np.array(steps)
    np.savez_compressed(os.path.join(HERE, 'pl.npz'), **out)
    print('pl.npz', len(out))


def gen_sg3d():
    """StyleGAN3-style discriminator and conv2d_resample through the reference's own code (CPU, fp32)."""
    from implementations.StyleGAN3 import model as sg3
    from thirdparty.stylegan3_ops.ops import conv2d_resample as ref_cr
    out = {}
    torch.manual_seed(99)
    # conv2d_resample cases: (name, ci, co, k, down, padding, use filter)
    cases = [('same3', 6, 8, 3, 1, 1, False), ('down3', 8, 8, 3, 2, 1, True), ('down1', 6, 12, 1, 2, 0, True),
             ('valid3', 4, 4, 3, 1, 0, False), ('same1', 3, 8, 1, 1, 0, False), ('asym3', 4, 6, 3, 1, [2, 0, 1, 1], False)]
    out['cr.cases'] = np.array([repr(c) for c in cases])
    f = ref_up.setup_filter([1, 3, 3, 1])
    out['cr.f'] = A(f)
    for name, ci, co, k, down, pad, use_f in cases:
        x = torch.randn(2, ci, 18, 14, requires_grad=True)
        w = torch.randn(co, ci, k, k, requires_grad=True)
        y = ref_cr.conv2d_resample(x, w, f if use_f else None, 1, down, pad)
        gy = torch.randn_like(y)
        gx, gw = torch.autograd.grad(y, (x, w), gy)
        for key, t in (('x', x), ('w', w), ('y', y), ('gy', gy), ('gx', gx), ('gw', gw)):
            out[f'cr.{name}.{key}'] = A(t)
    # the discriminator: 32 px, channels 8 -> 32, batch 8 (two minibatch-stddev groups of 4)
    cfg = dict(image_size=32, in_channels=3, channels=8, max_channels=32, mbsd_group_size=4, mbsd_channels=1)
    out['cfg'] = np.array(repr(cfg))
    D = sg3.Discriminator(**cfg)
    for k, v in D.state_dict().items():
        out['D0.' + k] = A(v)
    real = torch.rand(8, 3, 32, 32) * 2 - 1
    fake = torch.rand(8, 3, 32, 32) * 2 - 1
    out['real'] = A(real); out['fake'] = A(fake)
    loss = NonSaturatingLoss()
    lr, lf = D(real), D(fake)
    out['logits_real'] = A(lr); out['logits_fake'] = A(lf)
    d_loss = loss.d_loss(lr, lf)
    out['d_loss'] = A(d_loss)
    dg = torch.autograd.grad(d_loss, list(D.parameters()), allow_unused=True)
    for (n_, p), g_ in zip(D.named_parameters(), dg):
        out['dgrad.' + n_] = A(g_) if g_ is not None else np.zeros(p.shape, np.float32)
    r1 = r1_regularizer()(real, D, None)
    out['r1'] = A(r1)
    r1g = torch.autograd.grad(r1, list(D.parameters()), allow_unused=True)
    for (n_, p), g_ in zip(D.named_parameters(), r1g):
        out['r1grad.' + n_] = A(g_) if g_ is not None else np.zeros(p.shape, np.float32)
        out['r1none.' + n_] = np.array(g_ is None)
    np.savez_compressed(os.path.join(HERE, 'sg3d.npz'), **out)
    print('sg3d.npz', len(out))


def gen_ada():
    """thirdparty/ada/augment.py AugmentPipe and nnutils/ada.py ADA through the reference's own code (CPU, fp32) with every
    random draw recorded: outputs, image gradients, an R1-pattern second-order gradient, and the p-update sequence."""
    from nnutils.ada import ADA
    from thirdparty.ada.augment import AugmentPipe
    out = {}
    torch.manual_seed(31337)
    full = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1)
    cases = [('full', full, 0.8, (4, 3, 32, 32)), ('geom', dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1), 1.0, (3, 3, 24, 40)),
             ('color', dict(brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1), 1.0, (4, 3, 16, 16)),
             ('gray', dict(xflip=1, scale=1, brightness=1, contrast=1, lumaflip=1), 1.0, (3, 1, 16, 16)),
             ('extra', dict(xint=1, imgfilter=1, noise=1, cutout=1), 1.0, (2, 3, 32, 32)),
             ('blit', dict(xflip=1, rotate90=1, xint=1), 1.0, (4, 3, 16, 16))]
    out['cases'] = np.array([repr((n_, kw, p_, shape)) for n_, kw, p_, shape in cases])
    for name, kw, p_, shape in cases:
        pipe = AugmentPipe(**kw)
        pipe.p.copy_(torch.tensor(p_))
        x = (torch.rand(*shape) * 2 - 1).requires_grad_(True)
        with Recorder() as rec:
            y = pipe(x)
        gy = torch.randn_like(y)
        gx, = torch.autograd.grad(y, x, gy)
        out[f'{name}.x'] = A(x); out[f'{name}.y'] = A(y); out[f'{name}.gy'] = A(gy); out[f'{name}.gx'] = A(gx)
        out[f'{name}.n_draws'] = np.array(len(rec.items))
        for i, t in enumerate(rec.items):
            out[f'{name}.draw.{i}'] = A(t)
    # debug_percentile path (deterministic parameters)
    pipe = AugmentPipe(**full)
    x = torch.rand(2, 3, 32, 32) * 2 - 1
    with Recorder() as rec:
        y = pipe(x, debug_percentile=0.7)
    out['pct.x'] = A(x); out['pct.y'] = A(y); out['pct.n_draws'] = np.array(len(rec.items))
    for i, t in enumerate(rec.items):
        out[f'pct.draw.{i}'] = A(t)
    # ADA.update_p: 12 calls with interval 4
    ada = ADA(batch_size=8, interval=4, target_kimg=1, threshold=0.6)
    probs = torch.randn(12, 8) + 0.8
    ps = []
    for i in range(12):
        ada.update_p(probs[i])
        ps.append(float(ada.p))
    out['ada.probs'] = A(probs); out['ada.p'] = np.array(ps, np.float64)
    out['ada.state_keys'] = np.array(sorted(ada.state_dict().keys()))
    out['ada.Hz_geom'] = A(ada.Hz_geom); out['ada.Hz_fbank'] = A(ada.Hz_fbank)
    np.savez_compressed(os.path.join(HERE, 'ada.npz'), **out)
    print('ada.npz', len(out))


def gen_sg3g():
    """StyleGAN3 generator (implementations/StyleGAN3/model.py:32-380) and filtered_lrelu (thirdparty/stylegan3_ops/ops/
    filtered_lrelu.py, reference path) through the reference's own code (CPU, fp32): image, updated EMA buffers, all parameter
    gradients; stand-alone filtered_lrelu cases incl. gradients."""
    from implementations.StyleGAN3 import model as sg3
    from thirdparty.stylegan3_ops.ops import filtered_lrelu as ref_fl
    out = {}
    torch.manual_seed(4711)
    cfg = dict(image_size=32, latent_dim=32, num_layers=6, map_num_layers=2, channels=16, max_channels=32, style_dim=32)
    out['cfg'] = np.array(repr(cfg))
    G = sg3.Generator(**cfg)
    for p_ in G.parameters():                       # biases start at 0 / 1: move them so their gradients are exercised
        if p_.ndim == 1:
            p_.data.add_(torch.randn_like(p_) * 0.1)
    for k, v in G.state_dict().items():
        out['G0.' + k] = A(v)
    z = torch.randn(3, 32)
    out['z'] = A(z)
    G.train()
    img = G(z)
    gy = torch.randn_like(img)
    out['image'] = A(img); out['gy'] = A(gy)
    grads = torch.autograd.grad(img, list(G.parameters()), gy, allow_unused=True)
    for (n_, p_), g_ in zip(G.named_parameters(), grads):
        out['grad.' + n_] = A(g_) if g_ is not None else np.zeros(p_.shape, np.float32)
    for k, v in G.state_dict().items():             # buffers after one training-mode forward (magnitude EMAs, w_avg)
        if 'ema' in k or 'w_avg' in k:
            out['G1.' + k] = A(v)
    G.eval()
    out['image_eval_psi07'] = A(G(z, truncation_psi=0.7))
    # filtered_lrelu cases: (name, channels, hw, up, down, fu taps, fd (taps | 'radial'), padding, gain, slope, clamp)
    cases = [('u2d2', 4, 10, 2, 2, 12, 12, [9, 8, 9, 8], 2 ** 0.5, 0.2, 256), ('u4d2', 3, 16, 4, 2, 24, 12, [-6, -9, -6, -9], 2 ** 0.5, 0.2, 256),
             ('u2d2r', 4, 10, 2, 2, 12, 'radial', [9, 8, 9, 8], 2 ** 0.5, 0.2, 1.5), ('u1d1', 5, 8, 1, 1, 1, 1, 0, 1.0, 1.0, None),
             ('u2d1', 3, 8, 2, 1, 8, 1, [3, 4, 3, 4], 1.3, 0.1, None)]
    out['fl.cases'] = np.array([repr(c) for c in cases])
    for name, ch, hw, up, down, fu_t, fd_t, pad, gain, slope, clamp in cases:
        fu = sg3.design_filter(fu_t, 2.0, 2.0, 16.0) if fu_t > 1 else None
        fd = (sg3.design_filter(12, 2.0, 2.0, 16.0, radial=True) if fd_t == 'radial' else (sg3.design_filter(fd_t, 2.0, 2.0, 16.0) if fd_t > 1 else None))
        x = torch.randn(2, ch, hw, hw + 3, requires_grad=True)
        b = torch.randn(ch, requires_grad=True)
        y = ref_fl.filtered_lrelu(x, fu, fd, b, up, down, pad, gain, slope, clamp, impl='ref')
        g_ = torch.randn_like(y)
        gx, gb = torch.autograd.grad(y, (x, b), g_)
        for key, t in (('x', x), ('b', b), ('y', y), ('gy', g_), ('gx', gx), ('gb', gb)):
            out[f'fl.{name}.{key}'] = A(t)
        if fu is not None:
            out[f'fl.{name}.fu'] = A(fu)
        if fd is not None:
            out[f'fl.{name}.fd'] = A(fd)
    np.savez_compressed(os.path.join(HERE, 'sg3g.npz'), **out)
    print('sg3g.npz', len(out))


def gen_resample():
    """conv2d_resample up-sampling / grouped branches and conv2d_gradfix.conv_transpose2d through the reference's own code
    (thirdparty/stylegan3_ops/ops/conv2d_resample.py:40-141, conv2d_gradfix.py:29-46; CPU, fp32), incl. gradients."""
    import torch.nn.functional as F
    from thirdparty.stylegan3_ops.ops import conv2d_gradfix as ref_gf
    from thirdparty.stylegan3_ops.ops import conv2d_resample as ref_cr
    out = {}
    g = torch.Generator().manual_seed(2024)
    rn = lambda *s_: torch.randn(*s_, generator=g)
    f4 = ref_up.setup_filter([1, 3, 3, 1])
    f6 = ref_up.setup_filter([1, 5, 10, 10, 5, 1])
    out['f4'] = A(f4); out['f6'] = A(f6)
    # (name, ci, co, k, up, down, padding, filter, groups, flip_weight)
    cases = [('up3', 6, 8, 3, 2, 1, 1, 'f4', 1, True), ('up1', 6, 8, 1, 2, 1, 0, 'f4', 1, True), ('up3nf', 4, 6, 3, 2, 1, 1, 'f4', 1, False),
             ('updown3', 6, 4, 3, 2, 2, 1, 'f4', 1, True), ('up3f6', 4, 4, 3, 2, 1, 1, 'f6', 1, True), ('up4', 4, 4, 3, 4, 1, 1, 'f6', 1, True),
             ('grp3', 8, 12, 3, 1, 1, 1, None, 2, True), ('grpup3', 8, 8, 3, 2, 1, 1, 'f4', 4, True), ('grpdown3', 8, 8, 3, 1, 2, 1, 'f4', 2, True),
             ('asymup', 4, 6, 3, 2, 1, [2, 1, 0, 3], 'f4', 1, True)]
    out['cr.cases'] = np.array([repr(c) for c in cases])
    for name, ci, co, k, up, down, pad, fname, groups, flip_w in cases:
        x = rn(2, ci, 10, 7).requires_grad_(True)
        w = rn(co, ci // groups, k, k).requires_grad_(True)
        f = dict(f4=f4, f6=f6).get(fname)
        y = ref_cr.conv2d_resample(x, w, f, up, down, pad, groups, flip_w)
        gy = rn(*y.shape)
        gx, gw = torch.autograd.grad(y, (x, w), gy)
        for key, t in (('x', x), ('w', w), ('y', y), ('gy', gy), ('gx', gx), ('gw', gw)):
            out[f'cr.{name}.{key}'] = A(t)
    # conv_transpose2d (name, ci, co, k, stride, padding, output_padding, groups)
    tcases = [('t3s1', 6, 4, 3, 1, 1, 0, 1), ('t3s2', 6, 4, 3, 2, 1, 1, 1), ('t3s2p0', 4, 4, 3, 2, 0, 0, 1), ('t1s2', 4, 6, 1, 2, 0, 1, 1),
              ('t3s3', 4, 4, 3, 3, 2, 1, 1), ('t3s2g2', 8, 6, 3, 2, 1, 0, 2)]
    out['ct.cases'] = np.array([repr(c) for c in tcases])
    for name, ci, co, k, stride, pad, opad, groups in tcases:
        x = rn(2, ci, 6, 9).requires_grad_(True)
        w = rn(ci, co // groups, k, k).requires_grad_(True)
        b = rn(co).requires_grad_(True)
        y = ref_gf.conv_transpose2d(x, w, b, stride=stride, padding=pad, output_padding=opad, groups=groups)
        gy = rn(*y.shape)
        gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
        for key, t in (('x', x), ('w', w), ('b', b), ('y', y), ('gy', gy), ('gx', gx), ('gw', gw), ('gb', gb)):
            out[f'ct.{name}.{key}'] = A(t)
    np.savez_compressed(os.path.join(HERE, 'resample.npz'), **out)
    print('resample.npz', len(out))


if __name__ == '__main__':
    if len(sys.argv) > 1:                                         # regenerate only the named files
        for name in sys.argv[1:]:
            dict(ops=gen_ops, modules=gen_modules, model=gen_model, pl=gen_pl, sg3d=gen_sg3d, resample=gen_resample, ada=gen_ada, sg3g=gen_sg3g)[name]()
        sys.exit(0)
    gen_ops()
    gen_modules()
    gen_model()
    gen_pl()
    gen_sg3d()
    gen_resample()
    gen_ada()
    gen_sg3g()
    for f in ('ops.npz', 'modules.npz', 'model.npz', 'pl.npz', 'sg3d.npz', 'resample.npz', 'ada.npz', 'sg3g.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')
