"""GPU parity at model / step level: the product Generator, Discriminator and Trainer (every conv, resample,
bias_act, mbstd and optimizer kernel reached through the C ABI) against the reference's golden vectors.
The bar is north_star's: 1e-3 relative (to the tensor's scale), fp32."""
import ast

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BAR = 1e-3


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().float().cpu().numpy()


def _models(g_model):
    from animeface_b200.model import Discriminator, Generator
    c = ast.literal_eval(str(g_model['cfg']))
    G = Generator(c['image_size'], c['image_channels'], c['style_dim'], c['channels'], c['max_channels'],
                  c['block_num_conv'], c['map_num_layers'], True, 0.01)
    D = Discriminator(c['image_size'], c['image_channels'], c['channels'], c['max_channels'], c['block_num_conv'], c['mbsd_groups'])
    G.load_state_dict({k: torch.from_numpy(v) for k, v in g_model.sub('G0.').items()})
    D.load_state_dict({k: torch.from_numpy(v) for k, v in g_model.sub('D0.').items()})
    return c, G.to(DEV), D.to(DEV)


def _close(a, ref, tol, what):
    ref = np.asarray(ref)
    if np.abs(ref).max() < 1e-10:
        assert np.abs(N(a)).max() < 1e-7 if a is not None else True, what
    else:
        assert a is not None, what
        assert rel_err(N(a), ref) < tol, (what, rel_err(N(a), ref))


def test_forward_images_logits_and_gradients(g_model):
    from animeface_b200.model import supplied_noise
    from animeface_b200.nnutils.loss import NonSaturatingLoss, r1_regularizer
    c, G, D = _models(g_model)
    z, real = T(g_model['z']), T(g_model['real'])
    noises = [T(g_model[f'fwd.noise.{i}']) for i in range(int(g_model['fwd.n_noise']))]
    with supplied_noise(noises) as q:
        image, style = G(z)
        assert q.remaining == 0
    assert image.shape == g_model['fwd.image'].shape
    _close(style, g_model['fwd.style'], BAR, 'style')
    _close(image, g_model['fwd.image'], BAR, 'image')
    lf, lr = D(image), D(real)
    _close(lf, g_model['fwd.logits_fake'], BAR, 'logits_fake')
    _close(lr, g_model['fwd.logits_real'], BAR, 'logits_real')
    loss = NonSaturatingLoss()
    g_loss = loss.g_loss(lf)
    assert abs(float(g_loss) - float(g_model['g_loss'])) < BAR * abs(float(g_model['g_loss']))
    names = [n for n, _ in G.named_parameters()]
    gg = torch.autograd.grad(g_loss, list(G.parameters()), retain_graph=True, allow_unused=True)
    for n, gr in zip(names, gg):
        _close(gr, g_model['ggrad.' + n], BAR, 'ggrad.' + n)
    d_loss = loss.d_loss(lr, D(image.detach()))
    assert abs(float(d_loss) - float(g_model['d_loss'])) < BAR * abs(float(g_model['d_loss']))
    dn = [n for n, _ in D.named_parameters()]
    dg = torch.autograd.grad(d_loss, list(D.parameters()), allow_unused=True)
    for n, gr in zip(dn, dg):
        _close(gr, g_model['dgrad.' + n], BAR, 'dgrad.' + n)
    # R1: second-order autograd through every D kernel
    r1 = r1_regularizer()(real, D, None)
    assert abs(float(r1) - float(g_model['r1'])) < BAR * abs(float(g_model['r1']))
    r1g = torch.autograd.grad(r1, list(D.parameters()), allow_unused=True)
    for n, gr in zip(dn, r1g):
        _close(gr, g_model['r1grad.' + n], BAR, 'r1grad.' + n)


def test_three_step_trajectory(g_model):
    """Trainer.step x3 (step 2 is an R1 step, d_k=2) replaying the reference's random draws."""
    from animeface_b200 import rng
    from animeface_b200.train import TrainConfig, Trainer, build_optimizers
    from animeface_b200.model import Generator
    c, G, D = _models(g_model)
    cfg = TrainConfig(image_size=c['image_size'], style_dim=c['style_dim'], channels=c['channels'], max_channels=c['max_channels'],
                      block_num_conv=c['block_num_conv'], map_num_layers=c['map_num_layers'], mbsd_groups=c['mbsd_groups'],
                      batch_size=c['batch'], lr=c['lr'], beta1=c['betas'][0], beta2=c['betas'][1], d_k=c['d_k'], r1_lambda=c['r1_lambda'])
    G_ema = Generator(c['image_size'], c['image_channels'], c['style_dim'], c['channels'], c['max_channels'],
                      c['block_num_conv'], c['map_num_layers'], True, 0.01).to(DEV)
    G_ema.load_state_dict(G.state_dict())
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    tr = Trainer(cfg, G, G_ema, D, opt_g, opt_d)
    for it in range(int(g_model['traj.steps'])):
        draws = [T(g_model[f'traj.{it}.draw.{i}']) for i in range(int(g_model[f'traj.{it}.n_draws']))]
        with rng.replay(draws) as q:
            d_loss, g_loss, fake = tr.step(T(g_model[f'traj.{it}.real']))
            assert q.remaining == 0, 'draw order differs from the reference'
        rd, rg = float(g_model[f'traj.{it}.d_loss']), float(g_model[f'traj.{it}.g_loss'])
        assert abs(float(d_loss) - rd) < BAR * abs(rd), (it, float(d_loss), rd)
        assert abs(float(g_loss) - rg) < BAR * abs(rg), (it, float(g_loss), rg)
        _close(fake, g_model[f'traj.{it}.fake'], 2 * BAR, f'fake.{it}')
    for k, v in D.state_dict().items():
        _close(v, g_model['D3.' + k], 2 * BAR, 'D3.' + k)
    for k, v in G.state_dict().items():
        _close(v, g_model['G3.' + k], 2 * BAR, 'G3.' + k)
    for k, v in G_ema.state_dict().items():
        _close(v, g_model['E3.' + k], 2 * BAR, 'E3.' + k)


def test_full_size_step_runs_and_is_finite():
    """BASELINE config 2 (256 px, B=32 would need ~40 GB of activations; B=8 here keeps the test short):
    two steps incl. kernel-launch accounting; every loss finite, parameters changed, EMA follows."""
    from animeface_b200 import _lib
    from animeface_b200.train import TrainConfig, Trainer, build_models, build_optimizers
    torch.manual_seed(0)
    cfg = TrainConfig(batch_size=8, d_k=1)                  # d_k=1: step 1 is an R1 step
    G, G_ema, D = build_models(cfg, DEV)
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    tr = Trainer(cfg, G, G_ema, D, opt_g, opt_d)
    p0 = opt_g.flat_params.clone()
    before = _lib.launch_count()
    for _ in range(2):
        d_loss, g_loss, fake = tr.step(torch.rand(8, 3, 256, 256, device=DEV) * 2 - 1)
        assert torch.isfinite(d_loss) and torch.isfinite(g_loss) and torch.isfinite(fake).all()
    assert fake.shape == (8, 3, 256, 256) and float(fake.abs().max()) <= 1.0
    assert _lib.launch_count() - before > 300
    assert not torch.equal(p0, opt_g.flat_params)
    assert float((G_ema._sg2_flat - p0).abs().max()) > 0


def test_cuda_graph_trainer_matches_schedule_and_trains():
    """GraphedTrainer: eager -> capture -> replay for both step kinds; the replayed steps keep training (losses finite,
    parameters move, Adam step counters advance once per step for tensors that have a gradient)."""
    from animeface_b200.train import GraphedTrainer, TrainConfig, Trainer, build_models, build_optimizers
    torch.manual_seed(0)
    cfg = TrainConfig(image_size=32, style_dim=64, channels=8, max_channels=64, map_num_layers=2, batch_size=8, d_k=3)
    G, G_ema, D = build_models(cfg, DEV)
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    gt = GraphedTrainer(Trainer(cfg, G, G_ema, D, opt_g, opt_d))
    kinds, snaps = [], []
    for it in range(10):
        kinds.append(gt.t.is_r1_step())
        d_loss, g_loss, fake = gt.step(torch.rand(8, 3, 32, 32, device=DEV) * 2 - 1)
        assert torch.isfinite(d_loss) and torch.isfinite(g_loss) and torch.isfinite(fake).all(), it
        snaps.append(opt_d.flat_params.clone())
    assert kinds == [False, False, False, True, False, False, True, False, False, True]
    assert set(gt.graphs) == {(False, False), (True, False)}                      # both kinds were captured (2nd occurrence) ...
    assert gt.t.batches_done == 10
    for a, b in zip(snaps[:-1], snaps[1:]):
        assert not torch.equal(a, b)                            # ... and every replay really stepped the optimizer
    steps = opt_d._steps.cpu()
    assert int(steps.max()) == 10 and int(steps.min()) >= 7     # tensors without a gradient on the 3 R1 steps lag by 3


def test_path_length_penalty_value_and_second_order_gradients(g_pl):
    """a9: pl_penalty through the any-order modulated convolution vs the reference (value, per-sample norms via the
    penalty, every parameter gradient -- second order through conv / up2x / bias_act w.r.t. the style)."""
    from animeface_b200 import rng
    from animeface_b200.ops.conv2d import any_order_modconv
    from animeface_b200.train import pl_penalty
    c, G, _ = _models(g_pl)
    draws = [T(g_pl[f'eval.draw.{i}']) for i in range(int(g_pl['eval.n_draws']))]
    with rng.replay(draws) as q:
        with any_order_modconv():
            image, style = G(T(g_pl['z']))
        pl = pl_penalty(style, image, float(g_pl['pl_mean0']))
        assert q.remaining == 0
    _close(image, g_pl['eval.image'], BAR, 'image (any-order path)')
    assert abs(float(pl) - float(g_pl['eval.pl'])) < BAR * abs(float(g_pl['eval.pl'])), (float(pl), float(g_pl['eval.pl']))
    names = [n for n, _ in G.named_parameters()]
    pg = torch.autograd.grad(pl, list(G.parameters()), allow_unused=True)
    for n, gr in zip(names, pg):
        if bool(g_pl['plnone.' + n]):
            assert gr is None or float(gr.abs().max()) == 0, n
        else:
            _close(gr, g_pl['plgrad.' + n], BAR, 'plgrad.' + n)


def test_any_order_modconv_matches_fused(g_pl):
    """The composed (any-order) modulated convolution and the fused first-order ModConvFn give the same image and the
    same first-order parameter gradients."""
    from animeface_b200 import rng
    from animeface_b200.ops.conv2d import any_order_modconv
    c, G, _ = _models(g_pl)
    z = T(g_pl['z'])
    noises = [T(g_pl[f'eval.draw.{i}']) for i in range(int(g_pl['eval.n_draws']) - 1)]
    with rng.replay(list(noises)):
        a, _ = G(z)
    with rng.replay(list(noises)), any_order_modconv():
        b, _ = G(z)
    _close(b, N(a), 1e-5, 'image')
    w = torch.randn_like(a)
    ga = torch.autograd.grad((a * w).sum(), list(G.parameters()), allow_unused=True)
    gb = torch.autograd.grad((b * w).sum(), list(G.parameters()), allow_unused=True)
    for (n, _), x, y in zip(G.named_parameters(), ga, gb):
        assert (x is None) == (y is None), n
        if x is not None:
            _close(y, N(x), 1e-4, n)


def test_four_step_trajectory_with_path_length(g_pl):
    """Trainer.step x4 with pl_lambda > 0 (g_k=2: step 2 is a PL step; d_k=3: step 3 is an R1 step), reference draws."""
    from animeface_b200 import rng
    from animeface_b200.train import TrainConfig, Trainer, build_optimizers
    from animeface_b200.model import Generator
    c, G, D = _models(g_pl)
    cfg = TrainConfig(image_size=c['image_size'], style_dim=c['style_dim'], channels=c['channels'], max_channels=c['max_channels'],
                      block_num_conv=c['block_num_conv'], map_num_layers=c['map_num_layers'], mbsd_groups=c['mbsd_groups'],
                      batch_size=c['batch'], lr=c['lr'], beta1=c['betas'][0], beta2=c['betas'][1], d_k=c['d_k'], g_k=c['g_k'],
                      r1_lambda=c['r1_lambda'], pl_lambda=c['pl_lambda'])
    G_ema = Generator(c['image_size'], c['image_channels'], c['style_dim'], c['channels'], c['max_channels'],
                      c['block_num_conv'], c['map_num_layers'], True, 0.01).to(DEV)
    G_ema.load_state_dict(G.state_dict())
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    tr = Trainer(cfg, G, G_ema, D, opt_g, opt_d)
    for it in range(int(g_pl['traj.steps'])):
        draws = [T(g_pl[f'traj.{it}.draw.{i}']) for i in range(int(g_pl[f'traj.{it}.n_draws']))]
        assert tr.is_pl_step() == (it == 2) and tr.is_r1_step() == (it == 3)
        with rng.replay(draws) as q:
            d_loss, g_loss, fake = tr.step(T(g_pl[f'traj.{it}.real']))
            assert q.remaining == 0, 'draw order differs from the reference'
        rd, rg = float(g_pl[f'traj.{it}.d_loss']), float(g_pl[f'traj.{it}.g_loss'])
        assert abs(float(d_loss) - rd) < BAR * abs(rd), (it, float(d_loss), rd)
        assert abs(float(g_loss) - rg) < BAR * abs(rg), (it, float(g_loss), rg)
        rm = float(g_pl[f'traj.{it}.pl_mean'])
        assert abs(float(tr.pl_mean) - rm) <= BAR * abs(rm) + 1e-12, (it, float(tr.pl_mean), rm)
        _close(fake, g_pl[f'traj.{it}.fake'], 2 * BAR, f'fake.{it}')
    for k, v in D.state_dict().items():
        _close(v, g_pl['D4.' + k], 2 * BAR, 'D4.' + k)
    for k, v in G.state_dict().items():
        _close(v, g_pl['G4.' + k], 2 * BAR, 'G4.' + k)
    for k, v in G_ema.state_dict().items():
        _close(v, g_pl['E4.' + k], 2 * BAR, 'E4.' + k)


def test_amp_mode_step_matches_fp32_step(g_model):
    """The reference's DEFAULT run mode (MiniAccelerator(amp=True): fp16 autocast + GradScaler, implementations/StyleGAN2/
    utils.py:47,62-113,167) on this package: autocast regions, scaler.scale(loss).backward(), the scaler-aware R1
    ``calc_grad`` (nnutils/loss/penalty.py:11-26) and optimizer.step() through the scaler.  The kernels compute and store
    fp32 under autocast; what autocast still turns into half precision is the torch-side glue between them (the
    demodulation matmul, loss arithmetic), so two steps (the second an R1 step) land on the amp=False run to AMP's own
    precision -- 2e-3, against fp16's 1e-3 unit round-off -- not to fp32 round-off."""
    from animeface_b200 import rng
    from animeface_b200.diffaugment import DiffAugment
    from animeface_b200.nnutils import MiniAccelerator, update_ema
    from animeface_b200.nnutils.loss import NonSaturatingLoss, r1_regularizer
    from animeface_b200.train import TrainConfig, build_optimizers
    from animeface_b200.model import Generator

    def run(amp):
        c, G, D = _models(g_model)
        cfg = TrainConfig(image_size=c['image_size'], style_dim=c['style_dim'], channels=c['channels'], max_channels=c['max_channels'],
                          block_num_conv=c['block_num_conv'], map_num_layers=c['map_num_layers'], mbsd_groups=c['mbsd_groups'],
                          batch_size=c['batch'], lr=c['lr'], beta1=c['betas'][0], beta2=c['betas'][1], d_k=c['d_k'], r1_lambda=c['r1_lambda'])
        G_ema = Generator(c['image_size'], c['image_channels'], c['style_dim'], c['channels'], c['max_channels'],
                          c['block_num_conv'], c['map_num_layers'], True, 0.01).to(DEV)
        G_ema.load_state_dict(G.state_dict())
        opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
        acc = MiniAccelerator(amp=amp)
        opt_g, opt_d = acc.prepare(opt_g, opt_d)
        loss, r1 = NonSaturatingLoss(), r1_regularizer()
        out = []
        for it in (1, 2):                                           # the loop body of utils.py:53-116 (it = 2: R1 step, d_k = 2)
            draws = [T(g_model[f'traj.{it}.draw.{i}']) for i in range(int(g_model[f'traj.{it}.n_draws']))]
            real = T(g_model[f'traj.{it}.real'])
            with rng.replay(draws):
                z = rng.randn(real.size(0), cfg.style_dim, device=DEV)
                opt_g.zero_grad(); opt_d.zero_grad()
                with acc.autocast():
                    real_aug = DiffAugment(real, cfg.policy)
                    real_prob = D(real_aug)
                    fake, _ = G(z)
                    fake_aug = DiffAugment(fake, cfg.policy)
                    fake_prob = D(fake_aug.detach())
                    if it % cfg.d_k == 0:
                        d_loss = r1(real, D, acc.scaler) * cfg.r1_lambda * cfg.d_k
                    else:
                        d_loss = loss.d_loss(real_prob, fake_prob)
                acc.backward(d_loss)
                opt_d.step()
                z = rng.randn(real.size(0), cfg.style_dim, device=DEV)
                with acc.autocast():
                    fake, _ = G(z)
                    fake_prob = D(DiffAugment(fake, cfg.policy))
                    g_loss = loss.g_loss(fake_prob)
                acc.backward(g_loss)
                opt_g.step()
                update_ema(G, G_ema, cfg.ema_decay)
                acc.update()
            out.append((float(d_loss), float(g_loss)))
        return out, G._sg2_flat.clone(), D._sg2_flat.clone()

    la, ga, da = run(True)
    lf, gf, df = run(False)
    for (d1, g1), (d2, g2) in zip(la, lf):
        assert abs(d1 - d2) < 2e-3 * abs(d2) and abs(g1 - g2) < 2e-3 * abs(g2), (la, lf)
    assert rel_err(N(ga), N(gf)) < 2e-3 and rel_err(N(da), N(df)) < 2e-3, (rel_err(N(ga), N(gf)), rel_err(N(da), N(df)))
    # and the fp32 run is the reference's trajectory (steps 1 and 2 from the reference's step-1 weights are not in the golden;
    # the losses of a fresh start are): sanity that the loop above is the reference's loop
    assert all(np.isfinite(v) for pair in la + lf for v in pair)
