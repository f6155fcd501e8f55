"""CPU: the oracle (oracle/) against the golden vectors produced by the reference itself.
This is what pins the oracle (the reference has no tests of its own for this path)."""
import ast

import numpy as np
import pytest
import torch

from conftest import rel_err, up_cases
from oracle import ops_numpy as O
from oracle import sg2_torch as T

TOL = 2e-6   # fp32 reference vs fp32/fp64 restatement: rounding-order differences only


def test_upfirdn2d_cases(g_ops):
    for name, shape, taps, kw, wrap in up_cases(g_ops):
        x = g_ops[f'up.{name}.x']
        f = None if taps is None else g_ops[f'up.{name}.f']
        y = getattr(O, wrap)(x, f, **kw)
        assert y.shape == g_ops[f'up.{name}.y'].shape, name
        assert rel_err(y, g_ops[f'up.{name}.y']) < TOL, name


def test_setup_filter(g_ops):
    specs = [([1, 3, 3, 1], {}), ([1, 2, 1], dict(gain=4)), (list(range(1, 10)), {}),
             ([1, 3, 3, 1], dict(flip_filter=True, normalize=False)), ([1, 2, 3], dict(separable=True, gain=2))]
    for i, (spec, kw) in enumerate(specs):
        f = O.setup_filter(spec, **kw)
        assert f.shape == g_ops[f'sf.{i}'].shape
        assert rel_err(f, g_ops[f'sf.{i}']) < 1e-6


def test_bias_act_forward(g_ops):
    for act in [str(a) for a in g_ops['ba.acts']]:
        for variant, kw in (('plain', {}), ('clamp', dict(clamp=0.7, gain=1.3, alpha=0.3))):
            k = f'ba.{act}.{variant}'
            y = O.bias_act(g_ops[k + '.x'], g_ops[k + '.b'], dim=1, act=act, **kw)
            assert rel_err(y, g_ops[k + '.y']) < 5e-6, k


def test_sg2_resampling_identities(g_modules):
    for name in 'abc':
        x = g_modules[f'upblur.{name}.x']
        assert rel_err(O.blur3x3(O.bilinear_up2x(x)), g_modules[f'upblur.{name}.y']) < TOL
        assert rel_err(O.bilinear_up2x(x), g_modules[f'up.{name}.y']) < TOL
    assert rel_err(O.blur3x3(g_modules['blur.x']), g_modules['blur.y']) < TOL
    assert rel_err(O.avgpool2(g_modules['avg.x']), g_modules['avg.y']) < TOL
    # the analytic identities the kernels rely on: Blur2d == filter2d([1,2,1]), AvgPool == downsample2d([1,1])
    x = g_modules['blur.x']
    assert rel_err(O.filter2d(x, O.setup_filter([1, 2, 1])), g_modules['blur.y']) < TOL
    assert rel_err(O.downsample2d(x, O.setup_filter([1, 1])), g_modules['avg.y']) < TOL


def test_mbstd(g_modules):
    for name in ('g4', 'odd'):
        y = O.minibatch_stddev(g_modules[f'mbstd.{name}.x'], int(g_modules[f'mbstd.{name}.group']))
        assert rel_err(y, g_modules[f'mbstd.{name}.y']) < TOL
        yt = T.mbstd(torch.from_numpy(g_modules[f'mbstd.{name}.x']), int(g_modules[f'mbstd.{name}.group']))
        assert rel_err(yt.numpy(), g_modules[f'mbstd.{name}.y']) < TOL


def test_modulated_conv(g_modules):
    for name in ('k3', 'k1', 'k3b'):
        g = g_modules.sub(f'mod.{name}.')
        coef_a = 1.0 / np.sqrt(g['aw'].shape[1])
        s = (g['style'] * coef_a) @ g['aw'].T + g['ab'] + 1
        y = O.modulated_conv2d(g['x'], g['weight'], s, g['bias'], demod=bool(g['demod']))
        assert rel_err(y, g['y']) < 1e-5, name


def _sd(g, prefix, grad=False):
    return {k: torch.from_numpy(v.copy()).requires_grad_(grad and np.issubdtype(v.dtype, np.floating))
            for k, v in g.sub(prefix).items()}


def test_model_forward_and_grads(g_model):
    cfg = ast.literal_eval(str(g_model['cfg']))
    sd_g, sd_d = _sd(g_model, 'G0.', True), _sd(g_model, 'D0.', True)
    for k in sd_g:
        if k.endswith('.kernel'):
            sd_g[k].requires_grad_(False)
    z, real = torch.from_numpy(g_model['z']), torch.from_numpy(g_model['real'])
    noises = T.Draws([torch.from_numpy(g_model[f'fwd.noise.{i}']) for i in range(int(g_model['fwd.n_noise']))])
    image, style = T.generator(sd_g, z, T.ReplayDraws(noises))
    assert rel_err(image.detach().numpy(), g_model['fwd.image']) < 1e-5
    assert rel_err(style.detach().numpy(), g_model['fwd.style']) < 1e-5
    lf = T.discriminator(sd_d, image, cfg['mbsd_groups'])
    lr = T.discriminator(sd_d, real, cfg['mbsd_groups'])
    assert rel_err(lf.detach().numpy(), g_model['fwd.logits_fake']) < 1e-5
    assert rel_err(lr.detach().numpy(), g_model['fwd.logits_real']) < 1e-5
    g_loss = T.g_loss_ns(lf)
    names = [k for k, v in sd_g.items() if v.requires_grad]
    grads = torch.autograd.grad(g_loss, [sd_g[k] for k in names], retain_graph=True, allow_unused=True)
    for k, gr in zip(names, grads):
        ref = g_model['ggrad.' + k]
        if gr is None:
            assert np.abs(ref).max() == 0
        else:
            assert rel_err(gr.numpy(), ref) < 2e-4 or np.abs(ref).max() < 1e-12, k
    r1 = T.r1_penalty(sd_d, real, cfg['mbsd_groups'])
    assert abs(float(r1) - float(g_model['r1'])) / abs(float(g_model['r1'])) < 1e-5
    dn = list(sd_d.keys())
    r1g = torch.autograd.grad(r1, [sd_d[k] for k in dn], allow_unused=True)
    for k, gr in zip(dn, r1g):
        ref = g_model['r1grad.' + k]
        if gr is None:
            assert bool(g_model['r1none.' + k]) or np.abs(ref).max() == 0, k
        else:
            assert rel_err(gr.numpy(), ref) < 2e-4 or np.abs(ref).max() < 1e-12, k


def test_training_trajectory(g_model):
    """3 optimizer steps (the last one an R1 step) with the reference's recorded random draws."""
    cfg = ast.literal_eval(str(g_model['cfg']))
    sd_g, sd_d = _sd(g_model, 'G0.', True), _sd(g_model, 'D0.', True)
    for k in sd_g:
        if k.endswith('.kernel'):
            sd_g[k].requires_grad_(False)
    sd_e = {k: v.detach().clone() for k, v in sd_g.items()}
    scfg = T.StepConfig(latent_dim=cfg['style_dim'], r1_lambda=cfg['r1_lambda'], d_k=cfg['d_k'],
                        mbsd_groups=cfg['mbsd_groups'], lr=cfg['lr'], betas=cfg['betas'])
    g_lr, g_b, d_lr, d_b = T.adam_hparams(scfg)
    opt_g = torch.optim.Adam([v for v in sd_g.values() if v.requires_grad], lr=g_lr, betas=g_b)
    opt_d = torch.optim.Adam(list(sd_d.values()), lr=d_lr, betas=d_b)
    for it in range(int(g_model['traj.steps'])):
        draws = T.Draws([torch.from_numpy(g_model[f'traj.{it}.draw.{i}']) for i in range(int(g_model[f'traj.{it}.n_draws']))])
        real = torch.from_numpy(g_model[f'traj.{it}.real'])
        d_loss, g_loss, fake = T.train_step(sd_g, sd_d, sd_e, opt_g, opt_d, real, it, T.ReplayDraws(draws), scfg)
        assert draws.pos == len(draws.items)
        assert abs(float(d_loss) - float(g_model[f'traj.{it}.d_loss'])) <= 2e-4 * abs(float(g_model[f'traj.{it}.d_loss'])), it
        assert abs(float(g_loss) - float(g_model[f'traj.{it}.g_loss'])) <= 2e-4 * abs(float(g_model[f'traj.{it}.g_loss'])), it
        assert rel_err(fake.numpy(), g_model[f'traj.{it}.fake']) < 1e-3, it
    for k, v in sd_d.items():
        assert rel_err(v.detach().numpy(), g_model['D3.' + k]) < 1e-3 or np.abs(g_model['D3.' + k]).max() < 1e-6, k
    for k, v in sd_e.items():
        assert rel_err(v.detach().numpy(), g_model['E3.' + k]) < 1e-3 or np.abs(g_model['E3.' + k]).max() < 1e-6, k


def test_path_length_penalty(g_pl):
    """pl_penalty value, per-sample gradient norms and every second-order parameter gradient vs the reference."""
    sd_g = _sd(g_pl, 'G0.', True)
    for k in sd_g:
        if k.endswith('.kernel'):
            sd_g[k].requires_grad_(False)
    draws = T.Draws([torch.from_numpy(g_pl[f'eval.draw.{i}']) for i in range(int(g_pl['eval.n_draws']))])
    rng = T.ReplayDraws(draws)
    image, style = T.generator(sd_g, torch.from_numpy(g_pl['z']), rng)
    assert rel_err(image.detach().numpy(), g_pl['eval.image']) < 1e-5
    pl = T.pl_penalty(style, image, float(g_pl['pl_mean0']), rng)
    assert draws.pos == len(draws.items)
    assert abs(float(pl) - float(g_pl['eval.pl'])) < 1e-4 * abs(float(g_pl['eval.pl']))
    names = [k for k, v in sd_g.items() if v.requires_grad]
    grads = torch.autograd.grad(pl, [sd_g[k] for k in names], allow_unused=True)
    for k, gr in zip(names, grads):
        ref = g_pl['plgrad.' + k]
        if gr is None:
            assert bool(g_pl['plnone.' + k]) or np.abs(ref).max() == 0, k
        else:
            assert rel_err(gr.numpy(), ref) < 2e-4 or np.abs(ref).max() < 1e-12, k


def test_training_trajectory_with_path_length(g_pl):
    """4 optimizer steps with pl_lambda > 0 (step 2 a PL step, step 3 an R1 step), reference draws replayed."""
    cfg = ast.literal_eval(str(g_pl['cfg']))
    sd_g, sd_d = _sd(g_pl, 'G0.', True), _sd(g_pl, 'D0.', True)
    for k in sd_g:
        if k.endswith('.kernel'):
            sd_g[k].requires_grad_(False)
    sd_e = {k: v.detach().clone() for k, v in sd_g.items()}
    scfg = T.StepConfig(latent_dim=cfg['style_dim'], r1_lambda=cfg['r1_lambda'], d_k=cfg['d_k'], g_k=cfg['g_k'],
                        pl_lambda=cfg['pl_lambda'], mbsd_groups=cfg['mbsd_groups'], lr=cfg['lr'], betas=cfg['betas'])
    g_lr, g_b, d_lr, d_b = T.adam_hparams(scfg)
    opt_g = torch.optim.Adam([v for v in sd_g.values() if v.requires_grad], lr=g_lr, betas=g_b)
    opt_d = torch.optim.Adam(list(sd_d.values()), lr=d_lr, betas=d_b)
    state = dict(pl_mean=0.)
    for it in range(int(g_pl['traj.steps'])):
        draws = T.Draws([torch.from_numpy(g_pl[f'traj.{it}.draw.{i}']) for i in range(int(g_pl[f'traj.{it}.n_draws']))])
        real = torch.from_numpy(g_pl[f'traj.{it}.real'])
        d_loss, g_loss, fake = T.train_step(sd_g, sd_d, sd_e, opt_g, opt_d, real, it, T.ReplayDraws(draws), scfg, state)
        assert draws.pos == len(draws.items)
        assert abs(float(d_loss) - float(g_pl[f'traj.{it}.d_loss'])) <= 2e-4 * abs(float(g_pl[f'traj.{it}.d_loss'])), it
        assert abs(float(g_loss) - float(g_pl[f'traj.{it}.g_loss'])) <= 2e-4 * abs(float(g_pl[f'traj.{it}.g_loss'])), it
        assert abs(state['pl_mean'] - float(g_pl[f'traj.{it}.pl_mean'])) <= 2e-4 * abs(float(g_pl[f'traj.{it}.pl_mean'])) + 1e-12, it
        assert rel_err(fake.numpy(), g_pl[f'traj.{it}.fake']) < 1e-3, it
    for k, v in sd_g.items():
        assert rel_err(v.detach().numpy(), g_pl['G4.' + k]) < 1e-3 or np.abs(g_pl['G4.' + k]).max() < 1e-6, k
    for k, v in sd_e.items():
        assert rel_err(v.detach().numpy(), g_pl['E4.' + k]) < 1e-3 or np.abs(g_pl['E4.' + k]).max() < 1e-6, k


def test_sg3_discriminator_and_conv2d_resample(g_sg3d):
    """oracle/sg3d_torch.py against the reference's StyleGAN3-style discriminator (logits, D-loss and R1 gradients) and its
    conv2d_resample (outputs and gradients of the down-sampling / padding cases)."""
    from oracle import sg3d_torch as S
    g = g_sg3d
    f = torch.from_numpy(g['cr.f'])
    for case in g['cr.cases']:
        name, ci, co, k, down, pad, use_f = ast.literal_eval(str(case))
        x = torch.from_numpy(g[f'cr.{name}.x']).requires_grad_(True)
        w = torch.from_numpy(g[f'cr.{name}.w']).requires_grad_(True)
        y = S.conv2d_resample(x, w, f if use_f else None, down, pad)
        assert rel_err(y.detach().numpy(), g[f'cr.{name}.y']) < 1e-5, name
        gx, gw = torch.autograd.grad(y, (x, w), torch.from_numpy(g[f'cr.{name}.gy']))
        assert rel_err(gx.numpy(), g[f'cr.{name}.gx']) < 1e-5 and rel_err(gw.numpy(), g[f'cr.{name}.gw']) < 1e-5, name
    sd = {k: torch.from_numpy(v.copy()).requires_grad_(not k.endswith('down_filter')) for k, v in g.sub('D0.').items()}
    real, fake = torch.from_numpy(g['real']), torch.from_numpy(g['fake'])
    lr, lf = S.discriminator(sd, real), S.discriminator(sd, fake)
    assert rel_err(lr.detach().numpy(), g['logits_real']) < 1e-5 and rel_err(lf.detach().numpy(), g['logits_fake']) < 1e-5
    names = [k for k, v in sd.items() if v.requires_grad]
    d_loss = T.d_loss_ns(lr, lf)
    assert abs(float(d_loss) - float(g['d_loss'])) < 1e-5 * abs(float(g['d_loss']))
    dg = torch.autograd.grad(d_loss, [sd[k] for k in names], allow_unused=True)
    for k, gr in zip(names, dg):
        assert rel_err(gr.numpy(), g['dgrad.' + k]) < 2e-4 or np.abs(g['dgrad.' + k]).max() < 1e-12, k
    x = real.clone().requires_grad_(True)
    out = S.discriminator(sd, x)
    gx, = torch.autograd.grad(out, x, torch.ones_like(out), create_graph=True)
    r1 = gx.reshape(gx.shape[0], -1).norm(2, dim=1).pow(2).mean() / 2.
    assert abs(float(r1) - float(g['r1'])) < 1e-5 * abs(float(g['r1']))
    r1g = torch.autograd.grad(r1, [sd[k] for k in names], allow_unused=True)
    for k, gr in zip(names, r1g):
        ref = g['r1grad.' + k]
        if gr is None:
            assert bool(g['r1none.' + k]) or np.abs(ref).max() == 0, k
        else:
            assert rel_err(gr.numpy(), ref) < 2e-4 or np.abs(ref).max() < 1e-12, k


def test_oracle_conv2d_resample_up_and_grouped_cases(g_resample):
    """oracle/sg3d_torch.py (generic plan) against the reference's fast paths incl. gradients (tests/golden/resample.npz)."""
    import ast
    import torch
    from oracle import sg3d_torch as S
    g = g_resample
    filt = dict(f4=torch.from_numpy(g['f4']), f6=torch.from_numpy(g['f6']))
    for case in [ast.literal_eval(str(c)) for c in g['cr.cases']]:
        name, ci, co, k, up, down, pad, fname, groups, flip_w = case
        x = torch.from_numpy(g[f'cr.{name}.x']).requires_grad_(True)
        w = torch.from_numpy(g[f'cr.{name}.w']).requires_grad_(True)
        y = S.conv2d_resample_full(x, w, filt.get(fname), up, down, pad, groups, flip_w)
        assert rel_err(y.detach().numpy(), g[f'cr.{name}.y']) < 2e-6, name
        gx, gw = torch.autograd.grad(y, (x, w), torch.from_numpy(g[f'cr.{name}.gy']))
        assert rel_err(gx.numpy(), g[f'cr.{name}.gx']) < 2e-6 and rel_err(gw.numpy(), g[f'cr.{name}.gw']) < 2e-6, name
    for case in [ast.literal_eval(str(c)) for c in g['ct.cases']]:
        name, ci, co, k, stride, pad, opad, groups = case
        x, w, b = (torch.from_numpy(g[f'ct.{name}.{t}']) for t in 'xwb')
        y = S.conv_transpose2d(x, w, b, stride, pad, opad, groups)
        assert rel_err(y.numpy(), g[f'ct.{name}.y']) < 1e-6, name


def test_oracle_filtered_lrelu_cases():
    """oracle/sg3g_torch.py against the reference's filtered_lrelu incl. gradients (tests/golden/sg3g.npz)."""
    import ast
    import torch
    from conftest import Golden
    from oracle import sg3g_torch as S
    g = Golden('sg3g.npz')
    for case in [ast.literal_eval(str(c)) for c in g['fl.cases']]:
        name, ch, hw, up, down, fu_t, fd_t, pad, gain, slope, clamp = case
        fu = torch.from_numpy(g[f'fl.{name}.fu']) if f'fl.{name}.fu' in g else None
        fd = torch.from_numpy(g[f'fl.{name}.fd']) if f'fl.{name}.fd' in g else None
        x = torch.from_numpy(g[f'fl.{name}.x']).requires_grad_(True)
        b = torch.from_numpy(g[f'fl.{name}.b']).requires_grad_(True)
        y = S.filtered_lrelu(x, fu, fd, b, up, down, pad, gain, slope, clamp)
        assert rel_err(y.detach().numpy(), g[f'fl.{name}.y']) < 2e-6, name
        gx, gb = torch.autograd.grad(y, (x, b), torch.from_numpy(g[f'fl.{name}.gy']))
        assert rel_err(gx.numpy(), g[f'fl.{name}.gx']) < 2e-6 and rel_err(gb.numpy(), g[f'fl.{name}.gb']) < 2e-6, name
