"""GPU parity of the StyleGAN3-style discriminator path through the C ABI: conv2d_resample and the
Discriminator of animeface_b200/stylegan3.py against reference-generated goldens (tests/golden/sg3d.npz), and at the
full channel widths against the oracle (oracle/sg3d_torch.py) evaluated on the same GPU in fp32 and fp64."""
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


def test_conv2d_resample_golden(g_sg3d):
    from animeface_b200.ops import conv2d_resample as CR
    g = g_sg3d
    f = T(g['cr.f'])
    for case in g['cr.cases']:
        name, ci, co, k, down, pad, use_f = ast.literal_eval(str(case))
        for cl in (False, True):
            x = T(g[f'cr.{name}.x'])
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
            x.requires_grad_(True)
            w = T(g[f'cr.{name}.w']).requires_grad_(True)
            y = CR.conv2d_resample(x, w, f if use_f else None, 1, down, pad)
            assert y.shape == g[f'cr.{name}.y'].shape, name
            assert rel_err(N(y), g[f'cr.{name}.y']) < 1e-5, (name, cl)
            gx, gw = torch.autograd.grad(y, (x, w), T(g[f'cr.{name}.gy']))
            assert rel_err(N(gx), g[f'cr.{name}.gx']) < 1e-5 and rel_err(N(gw), g[f'cr.{name}.gw']) < 1e-5, (name, cl)


def _close(a, ref, tol, what):
    ref = np.asarray(ref)
    if np.abs(ref).max() < 1e-10:
        assert a is None or np.abs(N(a)).max() < 1e-7, what
    else:
        assert a is not None, what
        assert rel_err(N(a), ref) < tol, (what, rel_err(N(a), ref))


def test_sg3_discriminator_golden(g_sg3d):
    from animeface_b200.nnutils.loss import NonSaturatingLoss, r1_regularizer
    from animeface_b200.stylegan3 import Discriminator
    g = g_sg3d
    cfg = ast.literal_eval(str(g['cfg']))
    D = Discriminator(**cfg)
    assert sorted(D.state_dict()) == sorted(g.sub('D0.'))                 # same state_dict keys as the reference
    D.load_state_dict({k: torch.from_numpy(v) for k, v in g.sub('D0.').items()})
    D = D.to(DEV)
    real, fake = T(g['real']), T(g['fake'])
    lr, lf = D(real), D(fake)
    _close(lr, g['logits_real'], BAR, 'logits_real')
    _close(lf, g['logits_fake'], BAR, 'logits_fake')
    d_loss = NonSaturatingLoss().d_loss(lr, lf)
    assert abs(float(d_loss) - float(g['d_loss'])) < BAR * abs(float(g['d_loss']))
    names = [n for n, _ in D.named_parameters()]
    dg = torch.autograd.grad(d_loss, list(D.parameters()), allow_unused=True)
    for n, gr in zip(names, dg):
        _close(gr, g['dgrad.' + n], BAR, 'dgrad.' + n)
    r1 = r1_regularizer()(real, D, None)
    assert abs(float(r1) - float(g['r1'])) < BAR * abs(float(g['r1']))
    r1g = torch.autograd.grad(r1, list(D.parameters()), allow_unused=True)
    for n, gr in zip(names, r1g):
        if bool(g['r1none.' + n]):
            assert gr is None or float(gr.abs().max()) == 0, n
        else:
            _close(gr, g['r1grad.' + n], BAR, 'r1grad.' + n)


def test_sg3_discriminator_full_width_vs_oracle_on_gpu():
    """channels 64 .. 512 at 256 px (the layer shapes of BASELINE config 5's discriminator), B = 4."""
    from animeface_b200.nnutils.loss import NonSaturatingLoss
    from animeface_b200.stylegan3 import Discriminator
    from oracle import sg3d_torch as S
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    D = Discriminator(256).to(DEV)
    real = torch.rand(4, 3, 256, 256, device=DEV) * 2 - 1
    fake = torch.rand(4, 3, 256, 256, device=DEV) * 2 - 1
    loss = NonSaturatingLoss()
    names = [n for n, _ in D.named_parameters()]
    lr, lf = D(real), D(fake)
    dg = torch.autograd.grad(loss.d_loss(lr, lf), list(D.parameters()))

    def oracle(dtype):
        sd = {k: v.detach().to(dtype).requires_grad_(not k.endswith('down_filter')) for k, v in D.state_dict().items()}
        o_lr, o_lf = S.discriminator(sd, real.to(dtype)), S.discriminator(sd, fake.to(dtype))
        o_loss = torch.nn.functional.softplus(-o_lr).mean() + torch.nn.functional.softplus(o_lf).mean()
        return o_lr.detach(), o_lf.detach(), torch.autograd.grad(o_loss, [sd[n] for n in names])
    o32, o64 = oracle(torch.float32), oracle(torch.float64)
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    rows = [('logits_real', rel(lr, o64[0]), rel(o32[0], o64[0])), ('logits_fake', rel(lf, o64[1]), rel(o32[1], o64[1]))]
    rows += [('dgrad:' + n, rel(a, t), rel(b, t)) for n, a, b, t in zip(names, dg, o32[2], o64[2])]
    worst = max(rows, key=lambda r: r[1])
    print(f'\\nSG3-D full width vs fp64: worst {worst[0]} ours {worst[1]:.2e} | fp32 oracle {worst[2]:.2e}; '
          f'logits ours {rows[0][1]:.2e} | fp32 oracle {rows[0][2]:.2e}')
    bad = [r for r in rows if r[1] > max(BAR, 3 * r[2])]
    assert not bad, bad
