"""GPU parity: every libsg2b200 op, called through the C ABI, against the golden vectors from the reference
and against the CPU oracle on seeded inputs.  Tolerances (fp32): 1e-5 relative-to-scale for elementwise /
FIR ops, 1e-4 for convolution-class reductions (different summation order), far inside the 1e-3 bar."""
import numpy as np
import pytest
import torch

from conftest import rel_err, up_cases

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def T(a, grad=False):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV).requires_grad_(grad)


def N(t):
    return t.detach().double().cpu().numpy()


# ------------------------------------------------------------------------------------------ upfirdn2d
@pytest.mark.parametrize('fmt', ['contiguous', 'channels_last'])
def test_upfirdn2d_golden(g_ops, fmt):
    from animeface_b200.ops import upfirdn2d as U
    for name, shape, taps, kw, wrap in up_cases(g_ops):
        x = T(g_ops[f'up.{name}.x'])
        if fmt == 'channels_last':
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        f = None if taps is None else T(g_ops[f'up.{name}.f'])
        y = getattr(U, wrap)(x, f, **kw)
        assert rel_err(N(y), g_ops[f'up.{name}.y']) < 1e-5, name
        gx, = torch.autograd.grad(y, x, T(g_ops[f'up.{name}.gy']))
        assert rel_err(N(gx), g_ops[f'up.{name}.gx']) < 1e-5, name


def test_upfirdn2d_dtypes_and_big_nhwc():
    from animeface_b200.ops import upfirdn2d as U
    from oracle import ops_numpy as O
    rs = np.random.RandomState(0)
    x = rs.randn(2, 8, 33, 29).astype(np.float32)
    f = O.setup_filter([1, 3, 3, 1])
    for kw in (dict(up=2, padding=[2, 1, 2, 1], gain=4.0), dict(down=2, padding=[1, 1, 1, 1]), dict(padding=2)):
        ref = O.upfirdn2d(x, f, **kw)
        for dt, tol in ((torch.float32, 1e-5), (torch.float64, 1e-6), (torch.float16, 3e-3)):
            for cl in (False, True):
                xt = T(x).to(dt)
                if cl:
                    xt = xt.contiguous(memory_format=torch.channels_last)
                y = U.upfirdn2d(xt, T(f), **kw)
                assert y.dtype == dt and y.shape == ref.shape
                assert y.is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
                assert rel_err(N(y), ref) < tol, (kw, dt, cl)


def test_upfirdn2d_ring_path_variants():
    """The register-ring fast path (up = 1; 4x4 / 3x3 / 2x2 taps; down 1 or 2) against the oracle: asymmetric and
    negative padding, flipped non-symmetric filters, gains, ragged sizes, both layouts."""
    from animeface_b200.ops import upfirdn2d as U
    from oracle import ops_numpy as O
    rs = np.random.RandomState(1)
    x = rs.randn(3, 12, 37, 21).astype(np.float32)
    cases = [(rs.randn(4, 4), dict(padding=[2, 1, 0, 3], gain=1.7)), (rs.randn(4, 4), dict(down=2, padding=[1, 2, 2, 1], flip_filter=True)),
             (rs.randn(3, 3), dict(padding=[1, 1, 1, 1])), (rs.randn(3, 3), dict(padding=[-1, 2, 3, -2], flip_filter=True)),
             (rs.randn(2, 2), dict(down=2)), (rs.randn(4, 4), dict(down=2, padding=[-2, 5, 4, -1]))]
    for f, kw in cases:
        f = f.astype(np.float32)
        ref = O.upfirdn2d(x, f, **kw)
        for cl in (False, True):
            xt = T(x).contiguous(memory_format=torch.channels_last) if cl else T(x)
            y = U.upfirdn2d(xt, T(f), **kw)
            assert y.shape == ref.shape and rel_err(N(y), ref) < 1e-5, (f.shape, kw, cl)


def test_upfirdn2d_errors():
    from animeface_b200.ops import upfirdn2d as U
    x = torch.zeros(1, 1, 4, 4, device=DEV)
    with pytest.raises(RuntimeError):
        U.upfirdn2d(x, torch.ones(9, 9, device=DEV))                      # output < 1x1
    with pytest.raises(RuntimeError):
        U.upfirdn2d(x, torch.ones(2, 2, device=DEV, dtype=torch.float64))  # f must be float32
    with pytest.raises(RuntimeError):
        U.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))            # CPU tensors: no fallback


# ------------------------------------------------------------------------------------------- bias_act
def test_bias_act_golden(g_ops):
    from animeface_b200.ops.bias_act import bias_act
    for act in [str(a) for a in g_ops['ba.acts']]:
        for variant, kw in (('plain', {}), ('clamp', dict(clamp=0.7, gain=1.3, alpha=0.3))):
            k = f'ba.{act}.{variant}'
            x, b, gy = T(g_ops[k + '.x'], True), T(g_ops[k + '.b'], True), T(g_ops[k + '.gy'], True)
            y = bias_act(x, b, dim=1, act=act, **kw)
            assert rel_err(N(y), g_ops[k + '.y']) < 1e-5, k
            gx, gb = torch.autograd.grad(y, (x, b), gy, create_graph=True)
            assert rel_err(N(gx), g_ops[k + '.gx']) < 1e-5, k
            assert rel_err(N(gb), g_ops[k + '.gb']) < 1e-5, k
            d_gy, d_x = torch.autograd.grad(gx, (gy, x), T(g_ops[k + '.ggx']), allow_unused=True)
            assert rel_err(N(d_gy), g_ops[k + '.d_gy']) < 1e-5, k
            ref_dx = g_ops[k + '.d_x']
            if d_x is None:
                assert np.abs(ref_dx).max() == 0, k
            elif np.abs(ref_dx).max() > 0:
                assert rel_err(N(d_x), ref_dx) < 2e-5, k


def test_bias_act_layouts_and_dims():
    from animeface_b200.ops.bias_act import bias_act
    from oracle import ops_numpy as O
    rs = np.random.RandomState(1)
    x = rs.randn(4, 8, 6, 6).astype(np.float32)
    b = rs.randn(8).astype(np.float32)
    ref = O.bias_act(x, b, 1, 'lrelu')
    for cl in (False, True):
        xt = T(x).contiguous(memory_format=torch.channels_last) if cl else T(x)
        assert rel_err(N(bias_act(xt, T(b), act='lrelu')), ref) < 1e-6
    x2 = rs.randn(5, 7).astype(np.float32)                      # odd sizes -> scalar kernel, dim=1 on 2-D
    b2 = rs.randn(7).astype(np.float32)
    assert rel_err(N(bias_act(T(x2), T(b2), dim=1, act='tanh')), O.bias_act(x2, b2, 1, 'tanh')) < 1e-6
    for dt in (torch.float64, torch.float16):
        y = bias_act(T(x).to(dt), T(b).to(dt), act='sigmoid')
        assert y.dtype == dt and rel_err(N(y), O.bias_act(x, b, 1, 'sigmoid')) < (1e-6 if dt == torch.float64 else 2e-3)
    assert bias_act(torch.empty(0, 3, device=DEV), None).numel() == 0


# -------------------------------------------------------------------------------- StyleGAN2 resampling
def test_up2x_golden(g_modules):
    from animeface_b200.ops.resample import upsample2x_bilinear, upsample2x_blur
    for tag, fn in (('upblur', upsample2x_blur), ('up', upsample2x_bilinear)):
        for name in 'abc':        # a: C=4 (NHWC path), b: C=8 H=1 edge case, c: C=3 (NCHW path)
            x = T(g_modules[f'{tag}.{name}.x'], True)
            y = fn(x)
            assert rel_err(N(y), g_modules[f'{tag}.{name}.y']) < 1e-5, (tag, name)
            gx, = torch.autograd.grad(y, x, T(g_modules[f'{tag}.{name}.gy']))
            assert rel_err(N(gx), g_modules[f'{tag}.{name}.gx']) < 1e-5, (tag, name)


def test_up2x_vs_oracle_and_adjointness():
    from animeface_b200.ops.resample import Up2xAdjFn, upsample2x_blur
    from oracle import ops_numpy as O
    rs = np.random.RandomState(2)
    x = rs.randn(3, 16, 13, 9).astype(np.float32)
    y = upsample2x_blur(T(x))
    assert rel_err(N(y), O.blur3x3(O.bilinear_up2x(x))) < 1e-5
    # full-size (U1) property: <A x, g> == <x, A^T g>
    xb = torch.randn(32, 64, 128, 128, device=DEV).contiguous(memory_format=torch.channels_last)
    gb = torch.randn(32, 64, 256, 256, device=DEV).contiguous(memory_format=torch.channels_last)
    lhs = (upsample2x_blur(xb).double() * gb.double()).sum()
    rhs = (xb.double() * Up2xAdjFn.apply(gb, True).double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-6
    # DC gain: a constant image stays constant in the interior (filter sums to 1)
    c = upsample2x_blur(torch.ones(1, 4, 8, 8, device=DEV))
    assert float((c[:, :, 2:-2, 2:-2] - 1).abs().max()) < 1e-6


def test_avgpool2(g_modules):
    from animeface_b200.ops.resample import avgpool2
    from oracle import ops_numpy as O
    x = T(g_modules['avg.x'], True)
    y = avgpool2(x)
    assert rel_err(N(y), g_modules['avg.y']) < 1e-6
    rs = np.random.RandomState(3)
    a, b = rs.randn(2, 8, 10, 6).astype(np.float32), rs.randn(2, 8, 10, 6).astype(np.float32)
    at, bt = T(a, True), T(b, True)
    y = avgpool2(at, bt, 2 ** -0.5)
    ref = (O.avgpool2(a) + O.avgpool2(b)) / np.sqrt(2)
    assert rel_err(N(y), ref) < 1e-6
    gy = rs.randn(*ref.shape).astype(np.float32)
    ga, gb = torch.autograd.grad(y, (at, bt), T(gy))
    ref_g = np.repeat(np.repeat(gy, 2, 2), 2, 3) * (0.25 / np.sqrt(2))
    assert rel_err(N(ga), ref_g) < 1e-6 and rel_err(N(gb), ref_g) < 1e-6


# ---------------------------------------------------------------------------------------------- mbstd
@pytest.mark.parametrize('cl', [False, True])
def test_mbstd_golden(g_modules, cl):
    from animeface_b200.ops.mbstd import minibatch_stddev
    for name in ('g4', 'odd'):
        g = g_modules.sub(f'mbstd.{name}.')
        x = T(g['x'])
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        gy = T(g['gy'], True)
        y = minibatch_stddev(x, int(g['group']))
        assert rel_err(N(y), g['y']) < 1e-5
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        assert rel_err(N(gx), g['gx']) < 1e-5
        d_x, d_gy = torch.autograd.grad(gx, (x, gy), T(g['v']))
        assert rel_err(N(d_x), g['d_x']) < 1e-4
        assert rel_err(N(d_gy), g['d_gy']) < 1e-5


# ----------------------------------------------------------------------------------------- convolution
def test_modulated_conv_golden(g_modules):
    from animeface_b200.model import ModulatedConv2d
    for name in ('k3', 'k1', 'k3b'):
        g = g_modules.sub(f'mod.{name}.')
        co, ci, k, _ = g['weight'].shape
        m = ModulatedConv2d(ci, co, g['aw'].shape[1], k, demod=bool(g['demod'])).to(DEV)
        with torch.no_grad():
            m.weight.copy_(T(g['weight'])); m.bias.copy_(T(g['bias']))
            m.affine.layer.weight.copy_(T(g['aw'])); m.affine.layer.bias.copy_(T(g['ab']))
        x, w = T(g['x'], True), T(g['style'], True)
        y = m(x, w)
        assert rel_err(N(y), g['y']) < 1e-4, name
        grads = torch.autograd.grad(y, (x, w, m.weight, m.bias, m.affine.layer.weight, m.affine.layer.bias), T(g['gy']))
        for key, gr in zip(('gx', 'gstyle', 'gweight', 'gbias', 'gaw', 'gab'), grads):
            assert rel_err(N(gr), g[key]) < 2e-4, (name, key)


def test_dblock_golden_with_double_backward(g_modules):
    from animeface_b200.model import DBlock
    g = g_modules.sub('dblock.')
    blk = DBlock(8, 16, 2).to(DEV)
    names = [str(n) for n in g['names']]
    with torch.no_grad():
        for n, p in blk.named_parameters():
            p.copy_(T(g['p.' + n]))
    assert [n for n, _ in blk.named_parameters()] == names
    x = T(g['x'], True)
    y = blk(x)
    assert rel_err(N(y), g['y']) < 1e-4
    params = list(blk.parameters())
    grads = torch.autograd.grad(y, [x] + params, T(g['gy']), create_graph=True)
    assert rel_err(N(grads[0]), g['gx']) < 1e-4
    for n, gr in zip(names, grads[1:]):
        assert rel_err(N(gr), g['g.' + n]) < 1e-4, n
    pen = grads[0].pow(2).sum()
    g2 = torch.autograd.grad(pen, params, allow_unused=True)
    for n, gr in zip(names, g2):
        ref = g['g2.' + n]
        if gr is None:
            assert np.abs(ref).max() == 0, n
        else:
            assert rel_err(N(gr), ref) < 2e-4 or np.abs(ref).max() < 1e-10, n


@pytest.mark.parametrize('ci,co,k,hw', [(3, 32, 1, 17), (32, 64, 3, 16), (64, 64, 3, 9), (20, 36, 3, 7), (513, 64, 3, 4), (64, 3, 1, 12)])
def test_conv_family_vs_torch(ci, co, k, hw):
    """fwd / dgrad / wgrad against torch's fp32 conv (TF32 off) incl. odd channel counts and ragged tiles."""
    from animeface_b200.ops.conv2d import conv2d
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator(device=DEV).manual_seed(ci * 1000 + co)
    x = torch.randn(3, ci, hw, hw + 1, device=DEV, generator=gen, requires_grad=True)
    w = torch.randn(co, ci, k, k, device=DEV, generator=gen, requires_grad=True)
    coef = 0.37
    y = conv2d(x, w, coef)
    ref = torch.nn.functional.conv2d(x, w * coef, padding=k // 2)
    assert rel_err(N(y), N(ref)) < 1e-4
    gy = torch.randn(ref.shape, device=DEV, generator=gen)
    gx, gw = torch.autograd.grad(y, (x, w), gy)
    rx, rw = torch.autograd.grad(ref, (x, w), gy)
    assert rel_err(N(gx), N(rx)) < 1e-4 and rel_err(N(gw), N(rw)) < 1e-4


@pytest.mark.parametrize('n,ci,co,hw', [(5, 3, 32, 24), (2, 3, 64, 8), (3, 32, 3, 20), (4, 512, 3, 4), (2, 128, 3, 16), (3, 4, 8, 6), (2, 64, 1, 8),
                                         (2, 32, 3, 200), (1, 64, 3, 260)])
def test_thin_1x1_convs_vs_fp64(n, ci, co, hw):
    """RGB-side 1x1 layers (conv_thin.cu): fused forward incl. style / demod scales, bias, noise, lrelu and NCHW output,
    data gradient (the transposed layer is thin on the other side) and weight gradient with per-sample scales."""
    from animeface_b200.ops import conv2d as C
    import torch.nn.functional as F
    g = torch.Generator(device=DEV).manual_seed(ci * 7 + co)
    x = torch.randn(n, ci, hw, hw + 2, device=DEV, generator=g)
    w = torch.randn(co, ci, 1, 1, device=DEV, generator=g)
    s = torch.randn(n, ci, device=DEV, generator=g)
    d = torch.rand(n, co, device=DEV, generator=g) + 0.5
    b = torch.randn(co, device=DEV, generator=g)
    nz = torch.randn(n, 1, hw, hw + 2, device=DEV, generator=g)
    gy = torch.randn(n, co, hw, hw + 2, device=DEV, generator=g)
    coef = 0.11
    ref = F.leaky_relu(F.conv2d((x * s[:, :, None, None]).double(), (w * coef).double()) * d[:, :, None, None].double()
                       + b[None, :, None, None].double() + nz.double(), 0.2)
    for nchw in (False, True):
        y = C._conv_raw(x, w, coef, False, in_scale=s, out_scale=d, bias=b, noise=nz, slope=0.2, impl=1, out_nchw=nchw)
        assert rel_err(N(y), N(ref)) < 2e-6
    y0 = C._conv_raw(x, w, coef, False, impl=1)
    assert rel_err(N(y0), N(F.conv2d(x.double(), (w * coef).double()))) < 2e-6
    gx = C._conv_raw(gy, w, coef, True, impl=1)
    assert rel_err(N(gx), N(F.conv_transpose2d(gy.double(), (w * coef).double()))) < 2e-6
    wd = torch.zeros(co, ci, 1, 1, device=DEV, dtype=torch.float64, requires_grad=True)
    yr = F.conv2d((x * s[:, :, None, None]).double(), wd * coef) * d[:, :, None, None].double()
    ref_w, = torch.autograd.grad(yr, wd, gy.double())
    dw = C._wgrad_raw(x, gy, 1, coef, in_scale=s, out_scale=d, impl=1)
    assert rel_err(N(dw), N(ref_w)) < 5e-6
    dw0 = C._wgrad_raw(x, gy, 1, coef, impl=1)
    ref_w0, = torch.autograd.grad(F.conv2d(x.double(), wd * coef), wd, gy.double())
    assert rel_err(N(dw0), N(ref_w0)) < 5e-6


@pytest.mark.parametrize('n,ci,co,hw', [(8, 512, 1, 4), (3, 64, 2, 9)])
def test_thin_3x3_to_few_channels_vs_fp64(n, ci, co, hw):
    """3x3 layer with <= 4 output channels (the data gradient of the minibatch-stddev channel is 512 -> 1 @4^2)."""
    from animeface_b200.ops import conv2d as C
    import torch.nn.functional as F
    g = torch.Generator(device=DEV).manual_seed(ci + co)
    x = torch.randn(n, ci, hw, hw, device=DEV, generator=g)
    w = torch.randn(co, ci, 3, 3, device=DEV, generator=g)
    y = C._conv_raw(x, w, 0.2, False, impl=1)
    assert rel_err(N(y), N(F.conv2d(x.double(), (w * 0.2).double(), padding=1))) < 2e-6
    wt = torch.randn(ci, co, 3, 3, device=DEV, generator=g)          # as a data gradient: gy has ci channels
    gx = C._conv_raw(x, wt, 0.2, True, impl=1)
    assert rel_err(N(gx), N(F.conv_transpose2d(x.double(), (wt * 0.2).double(), padding=1))) < 2e-6


def test_conv_linearity_full_size():
    """Full-size (D block 1, conv2: 64->64 @256^2, B=32 would be 537 MB/tensor; B=8 keeps it quick):
    conv(a x1 + b x2) == a conv(x1) + b conv(x2), and the kernel is deterministic."""
    from animeface_b200.ops.conv2d import conv2d
    w = torch.randn(64, 64, 3, 3, device=DEV)
    x1 = torch.randn(8, 64, 256, 256, device=DEV).contiguous(memory_format=torch.channels_last)
    x2 = torch.randn_like(x1)
    lhs = conv2d(2.0 * x1 - 0.5 * x2, w, 1 / 24.)
    y1, y2 = conv2d(x1, w, 1 / 24.), conv2d(x2, w, 1 / 24.)
    assert rel_err(N(lhs), N(2.0 * y1 - 0.5 * y2)) < 1e-5
    assert torch.equal(y1, conv2d(x1, w, 1 / 24.))


# ------------------------------------------------------------------------------------------ optimizer
def test_flat_adam_matches_torch_adam_and_skips_none_grads():
    from animeface_b200.nnutils import FlatAdam, update_ema
    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)).to(DEV)
    a, b, ema_a, ema_b = mk(), mk(), mk(), mk()
    b.load_state_dict(a.state_dict()); ema_a.load_state_dict(a.state_dict()); ema_b.load_state_dict(a.state_dict())
    oa = torch.optim.Adam(a.parameters(), lr=1e-2, betas=(0., 0.99))
    ob = FlatAdam(b.parameters(), lr=1e-2, betas=(0., 0.99), model=b, ema_model=ema_b)
    for it in range(6):
        x = torch.randn(4, 7, device=DEV)
        for net, opt in ((a, oa), (b, ob)):
            opt.zero_grad()
            h = net[1](net[0](x))
            loss = (net[2](h) if it % 3 != 2 else h).square().sum()      # every 3rd step: last layer gets no grad
            loss.backward()
            opt.step()
        update_ema(a, ema_a, 0.9); update_ema(b, ema_b, 0.9)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert rel_err(N(pb), N(pa)) < 1e-5
    for pa, pb in zip(ema_a.parameters(), ema_b.parameters()):
        assert rel_err(N(pb), N(pa)) < 1e-5


@pytest.mark.parametrize('n,ci,co,k,h,w', [(2, 64, 64, 3, 33, 41), (1, 32, 128, 3, 65, 37), (3, 64, 32, 1, 40, 52), (2, 128, 64, 3, 129, 129),
                                            (32, 512, 512, 3, 4, 4), (16, 256, 128, 3, 8, 8), (5, 64, 64, 1, 4, 8)])
def test_tcgen05_ragged_images_vs_fp64(n, ci, co, k, h, w):
    """Images that do not tile by 8 x 16 (halo kernels: masked edge tiles) / by 32-pixel row pieces (wgrad: zero-filled
    overhang) -- the blurred 257^2 / 129^2 inputs of the StyleGAN3-style discriminator's down-sampling convolutions."""
    from animeface_b200.ops import conv2d as C
    import torch.nn.functional as F
    g = torch.Generator(device=DEV).manual_seed(h * w + ci)
    x = torch.randn(n, ci, h, w, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(co, ci, k, k, device=DEV, generator=g)
    s = torch.randn(n, ci, device=DEV, generator=g)
    d = torch.rand(n, co, device=DEV, generator=g) + 0.5
    b = torch.randn(co, device=DEV, generator=g)
    nz = torch.randn(n, 1, h, w, device=DEV, generator=g)
    gy = torch.randn(n, co, h, w, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    coef = 0.05
    ref = F.leaky_relu(F.conv2d((x * s[:, :, None, None]).double(), (wt * coef).double(), padding=k // 2) * d[:, :, None, None].double()
                       + b[None, :, None, None].double() + nz.double(), 0.2)
    for impl, tol in ((5, 3e-6), (4, 5e-5)):
        y = C._conv_raw(x, wt, coef, False, in_scale=s, out_scale=d, bias=b, noise=nz, slope=0.2, impl=impl)
        assert rel_err(N(y), N(ref)) < tol, impl
    gx = C._conv_raw(gy, wt, coef, True, impl=4)
    assert rel_err(N(gx), N(F.conv_transpose2d(gy.double(), (wt * coef).double(), padding=k // 2))) < 5e-5
    if w >= 32:
        wd = torch.zeros(co, ci, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
        yr = F.conv2d((x * s[:, :, None, None]).double(), wd * coef, padding=k // 2) * d[:, :, None, None].double()
        ref_w, = torch.autograd.grad(yr, wd, gy.double())
        dw = C._wgrad_raw(x, gy, k, coef, in_scale=s, out_scale=d, impl=4)
        assert rel_err(N(dw), N(ref_w)) < 5e-5


# ------------------------------------------------------------------------- tcgen05 kernels, every variant
@pytest.mark.parametrize('impl,tol', [(4, 5e-5), (5, 3e-6)])
@pytest.mark.parametrize('n,ci,co,k,hw', [(2, 64, 64, 3, 16), (8, 32, 64, 3, 16), (3, 64, 128, 3, 32), (4, 128, 32, 1, 16), (1, 256, 256, 3, 16),
                                          (2, 64, 192, 3, 16), (2, 96, 96, 3, 32), (1, 128, 320, 1, 16)])
def test_tcgen05_conv_variants_vs_fp64(impl, tol, n, ci, co, k, hw):
    """impl 4 = bf16x3, 5 = fp16x3 + promotion (fp32-class) halo kernels: plain, fused epilogue, data gradient; output widths
    with several 64- / 32-wide tiles (192, 96, 320) included."""
    from animeface_b200.ops import conv2d as C
    import torch.nn.functional as F
    g = torch.Generator(device=DEV).manual_seed(impl * 100 + ci)
    x = torch.randn(n, ci, hw, hw, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn(co, ci, k, k, device=DEV, generator=g)
    s = torch.randn(n, ci, device=DEV, generator=g)
    d = torch.rand(n, co, device=DEV, generator=g) + 0.5
    b = torch.randn(co, device=DEV, generator=g)
    nz = torch.randn(n, 1, hw, hw, device=DEV, generator=g)
    gy = torch.randn(n, co, hw, hw, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    coef = 0.05
    ref = F.leaky_relu(F.conv2d((x * s[:, :, None, None]).double(), (w * coef).double(), padding=k // 2) * d[:, :, None, None].double()
                       + b[None, :, None, None].double() + nz.double(), 0.2)
    y = C._conv_raw(x, w, coef, False, in_scale=s, out_scale=d, bias=b, noise=nz, slope=0.2, impl=impl)
    assert rel_err(N(y), N(ref)) < tol
    y_nchw = C._conv_raw(x, w, coef, False, in_scale=s, out_scale=d, bias=b, noise=nz, slope=0.2, impl=impl, out_nchw=True)
    assert y_nchw.is_contiguous() and torch.equal(y_nchw, y.contiguous())
    ref_t = F.conv_transpose2d(gy.double(), (w * coef).double(), padding=k // 2)
    gx = C._conv_raw(gy, w, coef, True, impl=impl)
    assert rel_err(N(gx), N(ref_t)) < tol


@pytest.mark.parametrize('n,ci,co,k,hw', [(8, 32, 64, 3, 16), (2, 64, 32, 3, 64), (8, 128, 128, 1, 16), (5, 96, 160, 3, 8), (8, 512, 512, 3, 4),
                                          (3, 64, 64, 3, 32), (1, 32, 32, 3, 128), (2, 192, 96, 1, 32), (3, 64, 256, 3, 4)])
def test_tcgen05_wgrad_vs_fp64(n, ci, co, k, hw):
    from animeface_b200.ops import conv2d as C
    import torch.nn.functional as F
    g = torch.Generator(device=DEV).manual_seed(ci + co)
    x = torch.randn(n, ci, hw, hw, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(n, co, hw, hw, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    s = torch.randn(n, ci, device=DEV, generator=g)
    d = torch.rand(n, co, device=DEV, generator=g) + 0.5
    w = torch.zeros(co, ci, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    coef = 0.07
    yr = F.conv2d((x * s[:, :, None, None]).double(), w * coef, padding=k // 2) * d[:, :, None, None].double()
    ref, = torch.autograd.grad(yr, w, gy.double())
    dw = C._wgrad_raw(x, gy, k, coef, in_scale=s, out_scale=d, impl=4)
    assert rel_err(N(dw), N(ref)) < 5e-5
    dw1 = C._wgrad_raw(x, gy, k, coef, in_scale=s, out_scale=d, impl=1)
    assert rel_err(N(dw1), N(ref)) < 5e-6
