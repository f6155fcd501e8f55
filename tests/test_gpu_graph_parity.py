"""GPU parity of the path bench.py actually times: ``GraphedTrainer`` (CUDA-graph capture + replay of ``Trainer.step``)
against the eager ``Trainer.step`` on the same weights, batches and random draws, and against the reference-generated
trajectory (tests/golden/model.npz: implementations/StyleGAN2/utils.py:53-116 run on the CPU by make_golden.py)."""
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


def _golden_trainer(g_model):
    from animeface_b200.model import Discriminator, Generator
    from animeface_b200.train import TrainConfig, Trainer, build_optimizers
    c = ast.literal_eval(str(g_model['cfg']))
    mk = lambda: Generator(c['image_size'], c['image_channels'], c['style_dim'], c['channels'], c['max_channels'],
                           c['block_num_conv'], c['map_num_layers'], True, 0.01)
    G, G_ema = mk(), mk()
    D = Discriminator(c['image_size'], c['image_channels'], c['channels'], c['max_channels'], c['block_num_conv'], c['mbsd_groups'])
    G.load_state_dict({k: torch.from_numpy(v) for k, v in g_model.sub('G0.').items()})
    D.load_state_dict({k: torch.from_numpy(v) for k, v in g_model.sub('D0.').items()})
    G_ema.load_state_dict(G.state_dict())
    G, G_ema, D = G.to(DEV), G_ema.to(DEV), D.to(DEV)
    cfg = TrainConfig(image_size=c['image_size'], style_dim=c['style_dim'], channels=c['channels'], max_channels=c['max_channels'],
                      block_num_conv=c['block_num_conv'], map_num_layers=c['map_num_layers'], mbsd_groups=c['mbsd_groups'],
                      batch_size=c['batch'], lr=c['lr'], beta1=c['betas'][0], beta2=c['betas'][1], d_k=c['d_k'], r1_lambda=c['r1_lambda'])
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    return Trainer(cfg, G, G_ema, D, opt_g, opt_d)


def _run_golden(g_model, graphed):
    from animeface_b200 import rng
    from animeface_b200.train import GraphedTrainer
    tr = _golden_trainer(g_model)
    runner = GraphedTrainer(tr, eager_first=False) if graphed else tr       # every step captured at its first occurrence
    out = []
    for it in range(int(g_model['traj.steps'])):
        draws = [T(g_model[f'traj.{it}.draw.{i}']) for i in range(int(g_model[f'traj.{it}.n_draws']))]
        if graphed:
            runner.graphs.clear()           # the draws are static tensors baked into the capture: one graph per step
        with rng.replay(draws) as q:
            d_loss, g_loss, fake = runner.step(T(g_model[f'traj.{it}.real']))
            assert q.remaining == 0
        out.append((float(d_loss), float(g_loss), fake.clone()))
    torch.cuda.synchronize()
    return tr, out


def test_graph_replay_matches_eager_and_reference_trajectory(g_model):
    """3 steps of the golden trajectory (step 2 is an R1 step), each CAPTURED and REPLAYED, vs the eager run and the reference."""
    _run_golden(g_model, False)                                   # warms every kernel up (lazy module loading cannot be captured)
    tr_e, eager = _run_golden(g_model, False)
    tr_g, graph = _run_golden(g_model, True)
    assert tr_g.batches_done == tr_e.batches_done == 3
    worst = 0.0
    for it, ((de, ge, fe), (dg, gg, fg)) in enumerate(zip(eager, graph)):
        rd, rg = float(g_model[f'traj.{it}.d_loss']), float(g_model[f'traj.{it}.g_loss'])
        assert abs(dg - rd) < BAR * abs(rd) and abs(gg - rg) < BAR * abs(rg), (it, dg, rd, gg, rg)
        assert rel_err(N(fg), g_model[f'traj.{it}.fake']) < 2 * BAR
        worst = max(worst, abs(dg - de) / abs(de), abs(gg - ge) / abs(ge), rel_err(N(fg), N(fe)))
    for name, a, b in (('G', tr_g.G, tr_e.G), ('D', tr_g.D, tr_e.D), ('E', tr_g.G_ema, tr_e.G_ema)):
        sa, sb = a.state_dict(), b.state_dict()
        for k in sa:
            if np.abs(N(sb[k])).max() > 0:
                worst = max(worst, rel_err(N(sa[k]), N(sb[k])))
            assert rel_err(N(sa[k]), g_model[f'{name}3.{k}']) < 2 * BAR, (name, k)
    print(f'\ngraph replay vs eager, golden trajectory: worst relative deviation {worst:.2e}')
    assert worst <= 1e-6, worst


def _random_run(graphed, steps=7):
    """A model wide enough for the tcgen05 kernels (channels 32..512 at 64 px), device RNG: the eager and the graphed run
    start from the same seeds, so the philox streams -- and with them every latent, noise map and augmentation -- agree."""
    from animeface_b200.train import GraphedTrainer, TrainConfig, Trainer, build_models, build_optimizers
    torch.manual_seed(11)
    cfg = TrainConfig(image_size=64, batch_size=8, d_k=3)
    G, G_ema, D = build_models(cfg, DEV)
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    tr = Trainer(cfg, G, G_ema, D, opt_g, opt_d)
    runner = GraphedTrainer(tr) if graphed else tr
    gen = torch.Generator(device=DEV).manual_seed(5)
    reals = [torch.rand(8, 3, 64, 64, device=DEV, generator=gen) * 2 - 1 for _ in range(steps)]
    torch.manual_seed(12)
    losses = []
    for it in range(steps):
        d_loss, g_loss, fake = runner.step(reals[it])
        losses.append((float(d_loss), float(g_loss), fake.clone()))
    torch.cuda.synchronize()
    if graphed:
        assert set(runner.graphs) == {(False, False), (True, False)}
    return tr, losses


def test_graph_replay_matches_eager_tensor_core_widths():
    """7 steps (R1 on steps 3 and 6; from step 1 on the normal steps are graph replays, step 6 is an R1 replay)."""
    tr_e, eager = _random_run(False)
    tr_g, graph = _random_run(True)
    worst_loss = worst_fake = 0.0
    for (de, ge, fe), (dg, gg, fg) in zip(eager, graph):
        worst_loss = max(worst_loss, abs(dg - de) / abs(de), abs(gg - ge) / abs(ge))
        worst_fake = max(worst_fake, rel_err(N(fg), N(fe)))
    worst_w = 0.0
    for a, b in ((tr_g.G, tr_e.G), (tr_g.D, tr_e.D), (tr_g.G_ema, tr_e.G_ema)):
        worst_w = max(worst_w, rel_err(N(a._sg2_flat), N(b._sg2_flat)))
    print(f'\ngraph replay vs eager, 64 px full width, 7 steps: losses {worst_loss:.2e}  fake {worst_fake:.2e}  weights {worst_w:.2e}')
    assert worst_loss <= 1e-6 and worst_fake <= 1e-6 and worst_w <= 1e-6, (worst_loss, worst_fake, worst_w)
