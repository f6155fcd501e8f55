"""GPU parity of the ADA augmentation pipeline (animeface_b200/ada.py on csrc/ada.cu + upfirdn2d) against reference-generated
goldens (tests/golden/ada.npz: thirdparty/ada/augment.py:115-427 and nnutils/ada.py:5-36 run by make_golden.py), replaying
the reference's random draws: outputs, image gradients, second order (R1 pattern), the p-update heuristic."""
import ast

import numpy as np
import pytest
import torch

from conftest import Golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def g_ada():
    return Golden('ada.npz')


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().float().cpu().numpy()


def test_filters_and_state_dict_match_the_reference(g_ada):
    from animeface_b200.ada import ADA
    ada = ADA(batch_size=8, interval=4, target_kimg=1, threshold=0.6)
    assert sorted(ada.state_dict().keys()) == [str(k) for k in g_ada['ada.state_keys']]
    assert rel_err(N(ada.Hz_geom), g_ada['ada.Hz_geom']) < 1e-6 and rel_err(N(ada.Hz_fbank), g_ada['ada.Hz_fbank']) < 1e-6


def test_augment_pipe_outputs_and_gradients(g_ada):
    from animeface_b200 import rng
    from animeface_b200.ada import AugmentPipe
    for case in [ast.literal_eval(str(c)) for c in g_ada['cases']]:
        name, kw, p_, shape = case
        pipe = AugmentPipe(**kw).to(DEV)
        pipe.p.copy_(torch.tensor(p_))
        x = T(g_ada[f'{name}.x']).requires_grad_(True)
        draws = [T(g_ada[f'{name}.draw.{i}']) for i in range(int(g_ada[f'{name}.n_draws']))]
        with rng.replay(draws) as q:
            y = pipe(x)
            assert q.remaining == 0, (name, 'draw order differs from the reference')
        assert tuple(y.shape) == g_ada[f'{name}.y'].shape
        e = rel_err(N(y), g_ada[f'{name}.y'])
        assert e < 1e-4, (name, e)
        gx, = torch.autograd.grad(y, x, T(g_ada[f'{name}.gy']))
        eg = rel_err(N(gx), g_ada[f'{name}.gx'])
        assert eg < 1e-4, (name, 'gx', eg)


def test_debug_percentile_path(g_ada):
    from animeface_b200 import rng
    from animeface_b200.ada import AugmentPipe
    full = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1)
    pipe = AugmentPipe(**full).to(DEV)
    draws = [T(g_ada[f'pct.draw.{i}']) for i in range(int(g_ada['pct.n_draws']))]
    with rng.replay(draws) as q:
        y = pipe(T(g_ada['pct.x']), debug_percentile=0.7)
        assert q.remaining == 0
    assert rel_err(N(y), g_ada['pct.y']) < 1e-4


def test_second_order_through_the_pipeline(g_ada):
    """The pipeline is linear in the image: the gradient of (A^T gy . v) w.r.t. gy is A v -- checked against the forward."""
    from animeface_b200 import rng
    from animeface_b200.ada import AugmentPipe
    name, kw, p_, shape = ast.literal_eval(str(g_ada['cases'][0]))
    pipe = AugmentPipe(**kw).to(DEV)
    pipe.p.copy_(torch.tensor(p_))
    draws = lambda: [T(g_ada[f'{name}.draw.{i}']) for i in range(int(g_ada[f'{name}.n_draws']))]
    x = T(g_ada[f'{name}.x']).requires_grad_(True)
    with rng.replay(draws()):
        y = pipe(x)
    gy = T(g_ada[f'{name}.gy']).requires_grad_(True)
    gx, = torch.autograd.grad(y, x, gy, create_graph=True)
    v = torch.randn_like(gx)
    av, = torch.autograd.grad(gx, gy, v)
    with rng.replay(draws()):
        y0 = pipe(torch.zeros_like(x))                      # the affine offset (brightness)
    with rng.replay(draws()):
        yv = pipe(v)
    assert rel_err(N(av), N(yv - y0)) < 1e-4


def test_ada_p_update_sequence(g_ada):
    from animeface_b200.ada import ADA
    ada = ADA(batch_size=8, interval=4, target_kimg=1, threshold=0.6).to(DEV)
    probs = T(g_ada['ada.probs'])
    got = []
    for i in range(probs.shape[0]):
        ada.update_p(probs[i])
        got.append(float(ada.p))
    assert np.allclose(np.array(got), g_ada['ada.p'], rtol=1e-6, atol=1e-9)
