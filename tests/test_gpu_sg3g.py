"""GPU parity of filtered_lrelu and the StyleGAN3 generator mirror (animeface_b200/stylegan3.py) against
reference-generated goldens (tests/golden/sg3g.npz: implementations/StyleGAN3/model.py:32-380 and
thirdparty/stylegan3_ops/ops/filtered_lrelu.py run by make_golden.py)."""
import ast

import numpy as np
import pytest
import torch

from conftest import Golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def g():
    return Golden('sg3g.npz')


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().float().cpu().numpy()


def test_filtered_lrelu_cases(g):
    from animeface_b200.ops.filtered_lrelu import filtered_lrelu
    for case in [ast.literal_eval(str(c)) for c in g['fl.cases']]:
        name, ch, hw, up, down, fu_t, fd_t, pad, gain, slope, clamp = case
        fu = T(g[f'fl.{name}.fu']) if f'fl.{name}.fu' in g else None
        fd = T(g[f'fl.{name}.fd']) if f'fl.{name}.fd' in g else None
        x, b = T(g[f'fl.{name}.x']).requires_grad_(True), T(g[f'fl.{name}.b']).requires_grad_(True)
        y = filtered_lrelu(x, fu, fd, b, up, down, pad, gain, slope, clamp)
        assert tuple(y.shape) == g[f'fl.{name}.y'].shape, name
        assert rel_err(N(y), g[f'fl.{name}.y']) < 1e-5, (name, rel_err(N(y), g[f'fl.{name}.y']))
        gx, gb = torch.autograd.grad(y, (x, b), T(g[f'fl.{name}.gy']))
        assert rel_err(N(gx), g[f'fl.{name}.gx']) < 1e-5, (name, 'gx')
        assert rel_err(N(gb), g[f'fl.{name}.gb']) < 1e-5, (name, 'gb')


def test_generator_image_buffers_and_gradients(g):
    from animeface_b200.stylegan3 import Generator
    cfg = ast.literal_eval(str(g['cfg']))
    G = Generator(**cfg)
    sd = {k: torch.from_numpy(v) for k, v in g.sub('G0.').items()}
    assert sorted(sd) == sorted(G.state_dict().keys())
    G.load_state_dict(sd)
    G = G.to(DEV).train()
    z = T(g['z'])
    img = G(z)
    assert rel_err(N(img), g['image']) < 1e-3, rel_err(N(img), g['image'])
    grads = torch.autograd.grad(img, list(G.parameters()), T(g['gy']), allow_unused=True)
    for (n_, p), gr in zip(G.named_parameters(), grads):
        ref = g['grad.' + n_]
        if np.abs(ref).max() < 1e-12:
            assert gr is None or float(gr.abs().max()) < 1e-7, n_
        else:
            assert rel_err(N(gr), ref) < 1e-3, (n_, rel_err(N(gr), ref))
    for k, v in G.state_dict().items():
        if 'ema' in k or 'w_avg' in k:
            assert rel_err(N(v), g['G1.' + k]) < 1e-4, k
    G.eval()
    assert rel_err(N(G(z, truncation_psi=0.7)), g['image_eval_psi07']) < 1e-3
