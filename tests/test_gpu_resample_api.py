"""GPU parity of the op-API branches behind conv2d_resample / conv2d_gradfix (up-sampling, transposed, grouped convolutions)
against reference-generated goldens (tests/golden/resample.npz: thirdparty/stylegan3_ops/ops/conv2d_resample.py:40-141 and
conv2d_gradfix.py:29-46 run by make_golden.py), forward and first-order gradients, plus a second-order (R1-pattern) check
against the oracle."""
import ast

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 2e-5


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().float().cpu().numpy()


def test_conv2d_resample_up_down_grouped(g_resample):
    from animeface_b200.ops import conv2d_resample as CR
    g = g_resample
    filt = dict(f4=T(g['f4']), f6=T(g['f6']))
    for case in [ast.literal_eval(str(c)) for c in g['cr.cases']]:
        name, ci, co, k, up, down, pad, fname, groups, flip_w = case
        x, w = T(g[f'cr.{name}.x']).requires_grad_(True), T(g[f'cr.{name}.w']).requires_grad_(True)
        y = CR.conv2d_resample(x, w, filt.get(fname), up, down, pad, groups, flip_w)
        assert tuple(y.shape) == g[f'cr.{name}.y'].shape, (name, tuple(y.shape))
        assert rel_err(N(y), g[f'cr.{name}.y']) < TOL, (name, rel_err(N(y), g[f'cr.{name}.y']))
        gx, gw = torch.autograd.grad(y, (x, w), T(g[f'cr.{name}.gy']))
        assert rel_err(N(gx), g[f'cr.{name}.gx']) < TOL, (name, 'gx')
        assert rel_err(N(gw), g[f'cr.{name}.gw']) < TOL, (name, 'gw')


def test_conv_transpose2d(g_resample):
    from animeface_b200.ops import conv2d_gradfix as GF
    g = g_resample
    for case in [ast.literal_eval(str(c)) for c in g['ct.cases']]:
        name, ci, co, k, stride, pad, opad, groups = case
        x, w, b = (T(g[f'ct.{name}.{t}']).requires_grad_(True) for t in 'xwb')
        y = GF.conv_transpose2d(x, w, b, stride=stride, padding=pad, output_padding=opad, groups=groups)
        assert tuple(y.shape) == g[f'ct.{name}.y'].shape, (name, tuple(y.shape))
        assert rel_err(N(y), g[f'ct.{name}.y']) < TOL, (name, rel_err(N(y), g[f'ct.{name}.y']))
        gx, gw, gb = torch.autograd.grad(y, (x, w, b), T(g[f'ct.{name}.gy']))
        for key, t in (('gx', gx), ('gw', gw), ('gb', gb)):
            assert rel_err(N(t), g[f'ct.{name}.{key}']) < TOL, (name, key)


def test_up_branch_second_order_vs_oracle(g_resample):
    """R1 pattern through the transposed-convolution plan: d/dw of |d sum(y * gy)/dx|^2 vs the oracle's generic plan."""
    from animeface_b200.ops import conv2d_resample as CR
    from oracle import sg3d_torch as S
    g = g_resample
    f = T(g['f4'])
    x0, w0, gy = T(g['cr.up3.x']), T(g['cr.up3.w']), T(g['cr.up3.gy'])
    res = []
    for fn in (lambda x, w: CR.conv2d_resample(x, w, f, 2, 1, 1), lambda x, w: S.conv2d_resample_full(x, w, f, 2, 1, 1)):
        x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        gx, = torch.autograd.grad((fn(x, w) * gy).sum(), x, create_graph=True)
        res.append(torch.autograd.grad(gx.square().sum(), w)[0])
    assert rel_err(N(res[0]), N(res[1])) < 1e-4
