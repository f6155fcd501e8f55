"""CPU: host logic of the training loop that needs no kernel -- the lazy-regularisation schedule (which steps are R1 /
path-length steps, implementations/StyleGAN2/utils.py:71-73, 96-98), the step kinds the CUDA-graph trainer captures, and
the Adam hyper-parameter scaling of utils.py:208-218."""
import types

import pytest

from animeface_b200.train import GraphedTrainer, TrainConfig, Trainer, update_pl_mean


def _trainer(cfg):
    t = Trainer.__new__(Trainer)            # schedule methods only: no models, no device
    t.cfg, t.batches_done = cfg, 0
    return t


def test_lazy_regularisation_schedule_matches_reference_conditions():
    t = _trainer(TrainConfig(d_k=16, g_k=8, r1_lambda=10., pl_lambda=2.))
    for it in range(0, 70):
        assert t.is_r1_step(it) == (it % 16 == 0 and it != 0)
        assert t.is_pl_step(it) == (it % 8 == 0 and it != 0)
    off = _trainer(TrainConfig(r1_lambda=0., pl_lambda=0.))
    assert not any(off.is_r1_step(it) or off.is_pl_step(it) for it in range(64))


@pytest.mark.parametrize('cfg,expected', [
    (TrainConfig(), [(False, False), (True, False)]),
    (TrainConfig(pl_lambda=2.), [(False, False), (False, True), (True, True)]),
    (TrainConfig(pl_lambda=2., g_k=3, d_k=2), [(False, False), (True, False), (False, True), (True, True)]),
    (TrainConfig(r1_lambda=0.), [(False, False)]),
])
def test_graph_kinds_cover_the_schedule(cfg, expected):
    gt = GraphedTrainer(_trainer(cfg))
    kinds = gt.kinds()
    assert kinds[0] == (False, False) and sorted(kinds) == sorted(expected)
    seen = {(gt.t.is_r1_step(it), gt.t.is_pl_step(it)) for it in range(1, 500)}
    assert seen == set(kinds)


def test_adam_hparams_scaling():
    from animeface_b200.train import build_optimizers
    made = []

    class FakeAdam:
        def __init__(self, params, lr, betas, model=None, ema_model=None):
            made.append((lr, betas))
    import animeface_b200.train as T
    real, T.FlatAdam = T.FlatAdam, FakeAdam
    try:
        net = types.SimpleNamespace(parameters=lambda: [])
        build_optimizers(TrainConfig(lr=1e-3, beta1=0., beta2=0.99, d_k=16, g_k=8, r1_lambda=10., pl_lambda=0.), net, net, net)
        build_optimizers(TrainConfig(lr=1e-3, beta1=0., beta2=0.99, d_k=16, g_k=8, r1_lambda=10., pl_lambda=2.), net, net, net)
    finally:
        T.FlatAdam = real
    (g_lr, g_b), (d_lr, d_b), (g_lr2, g_b2), _ = made
    assert g_lr == 1e-3 and g_b == (0., 0.99)
    assert d_lr == pytest.approx(1e-3 * 16 / 17) and d_b[1] == pytest.approx(0.99 ** (16 / 17))
    assert g_lr2 == pytest.approx(1e-3 * 8 / 9) and g_b2[1] == pytest.approx(0.99 ** (8 / 9))


def test_update_pl_mean():
    assert update_pl_mean(0., 2.) == pytest.approx(0.02)
    assert update_pl_mean(1., 3., decay=0.5) == pytest.approx(2.)
