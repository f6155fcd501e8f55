"""GPU parity of the bf16 pair-planes kernels (csrc/planes.cu, conv_halo_pl.cu, wgrad_pl.cu) -- the first-order backward fast
path -- against plain torch fp32 (TF32 off), through the C ABI.  Integer-exact where the arithmetic is (the bf16 split), 1e-5
class for the bf16x3 convolutions, bit-identical run to run for every reduction (deterministic two-pass split-K)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _cl(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).contiguous(memory_format=torch.channels_last)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _unplane(p):
    """[2][n,h,w,c] bf16 -> (hi, lo) as NCHW-shaped fp32 tensors."""
    return p[0].float().permute(0, 3, 1, 2), p[1].float().permute(0, 3, 1, 2)


def test_split_planes_is_the_exact_bf16_pair():
    from animeface_b200.ops import conv2d as C
    x = _cl(3, 64, 9, 7, seed=1, scale=3.0)
    s = torch.rand(3, 64, device=DEV) + 0.5
    for scale in (None, s):
        hi, lo = _unplane(C._split_planes(x, scale))
        v = x if scale is None else x * scale[:, :, None, None]
        want_hi = v.bfloat16().float()
        want_lo = (v - want_hi).bfloat16().float()
        assert torch.equal(hi, want_hi) and torch.equal(lo, want_lo)
        assert _rel(hi + lo, v) < 2.0 ** -16


@pytest.mark.parametrize('slope,with_d', [(0.2, False), (0.2, True), (None, False), (None, True)])
def test_bwd_prep_planes_matches_torch_and_is_deterministic(slope, with_d):
    from animeface_b200.ops import conv2d as C
    n, co, h, w = 4, 96, 12, 20
    gy, acc = _cl(n, co, h, w, seed=2), _cl(n, co, h, w, seed=3)
    d = (torch.rand(n, co, device=DEV) + 0.5) if with_d else None
    bias = torch.randn(1, co, 1, 1, device=DEV) * 0.3
    noise = torch.randn(n, 1, h, w, device=DEV)
    u = acc * (d[:, :, None, None] if with_d else 1.0) + bias + noise
    y = F.leaky_relu(u, slope) if slope is not None else u
    planes, gb, gd = C._bwd_prep_planes(gy, y, slope, noise=noise, bias=bias, d=d)
    gu = gy * torch.where(y > 0, 1.0, slope) if slope is not None else gy
    want = gu * (d[:, :, None, None] if with_d else 1.0)
    hi, lo = _unplane(planes)
    assert torch.equal(hi, want.bfloat16().float())
    assert _rel(hi + lo, want) < 2.0 ** -16
    assert _rel(gb, gu.sum((0, 2, 3))) < 1e-5
    if with_d:
        assert _rel(gd, (gu * acc).sum((2, 3))) < 2e-5
    p2, gb2, gd2 = C._bwd_prep_planes(gy, y, slope, noise=noise, bias=bias, d=d)
    assert torch.equal(planes, p2) and torch.equal(gb, gb2) and (gd is None or torch.equal(gd, gd2))


CONV_CASES = [  # n, cin (planes), cout, k, h, w
    (2, 64, 64, 3, 16, 16), (2, 64, 32, 3, 32, 32), (3, 128, 64, 3, 16, 32), (2, 64, 128, 1, 16, 16), (1, 256, 256, 3, 16, 16),
    (2, 64, 64, 3, 40, 33), (4, 512, 512, 3, 8, 8), (2, 128, 128, 3, 64, 64), (1, 64, 64, 1, 32, 48),
    (2, 32, 64, 3, 32, 32), (2, 32, 32, 3, 64, 64), (2, 96, 32, 3, 16, 16), (2, 32, 64, 1, 16, 16),
]


@pytest.mark.parametrize('n,cin,cout,k,h,w', CONV_CASES)
def test_data_gradient_conv_on_planes(n, cin, cout, k, h, w):
    """gx = conv_transpose2d(gy, w): the halo kernel fed by TMA from planes vs torch, and vs the fp32-operand halo kernel."""
    from animeface_b200.ops import conv2d as C
    _setup()
    if not C._planes_ok(n, h, w, cin, cout, k, False):
        pytest.skip('shape not taken by the planes kernel')
    gy = _cl(n, cin, h, w, seed=4)
    wt = torch.randn(cin, cout, k, k, device=DEV)                # forward weight [co=cin, ci=cout]: dgrad maps cin -> cout
    coef = 0.07
    ref = F.conv_transpose2d(gy, wt * coef, padding=k // 2)
    got = C._conv_planes(C._split_planes(gy), wt, coef, True)
    assert got.shape == ref.shape
    assert _rel(got, ref) < 2e-5, _rel(got, ref)
    old = C._conv_raw(gy, wt, coef, True)
    assert _rel(got, old) < 2e-5


WGRAD_CASES = [  # n, ci, co, k, h, w
    (2, 64, 64, 3, 16, 16), (2, 64, 128, 3, 32, 32), (3, 128, 64, 3, 8, 8), (2, 64, 192, 1, 16, 16), (1, 256, 256, 3, 16, 16),
    (2, 64, 64, 3, 40, 33), (4, 128, 512, 3, 8, 16), (2, 64, 64, 3, 64, 64), (2, 128, 128, 1, 32, 32), (5, 64, 64, 3, 8, 8),
    (2, 32, 64, 3, 32, 32), (2, 64, 32, 3, 32, 32), (2, 32, 32, 3, 64, 64), (2, 32, 64, 1, 16, 16), (2, 96, 160, 3, 16, 16),
]


@pytest.mark.parametrize('n,ci,co,k,h,w', WGRAD_CASES)
def test_weight_gradient_on_planes(n, ci, co, k, h, w):
    from animeface_b200.ops import conv2d as C
    _setup()
    if not C._planes_ok(n, h, w, ci, co, k, True):
        pytest.skip('shape not taken by the planes kernel')
    x, gy = _cl(n, ci, h, w, seed=5), _cl(n, co, h, w, seed=6)
    s = torch.rand(n, ci, device=DEV) + 0.5
    coef = 0.05
    wr = torch.zeros(co, ci, k, k, device=DEV, requires_grad=True)
    ref, = torch.autograd.grad(F.conv2d(x * s[:, :, None, None], wr * coef, padding=k // 2), wr, gy)
    xp, gyp = C._split_planes(x, s), C._split_planes(gy)
    got = C._wgrad_planes(xp, gyp, k, coef)
    assert _rel(got, ref) < 2e-5, _rel(got, ref)
    assert torch.equal(got, C._wgrad_planes(xp, gyp, k, coef))           # deterministic split-K


@pytest.mark.parametrize('n,ci,co,k,hw', [(2, 64, 64, 3, 32), (2, 32, 64, 3, 32), (4, 3, 32, 1, 64), (4, 32, 3, 1, 64), (2, 24, 40, 3, 12)])
def test_legacy_weight_gradient_kernels_are_deterministic(n, ci, co, k, hw):
    """fp32-operand wgrad kernels (tcgen05 with in-kernel transform, SIMT, thin): two-pass reduction through the workspace."""
    from animeface_b200.ops import conv2d as C
    _setup()
    x, gy = _cl(n, ci, hw, hw, seed=7), _cl(n, co, hw, hw, seed=8)
    wr = torch.zeros(co, ci, k, k, device=DEV, requires_grad=True)
    ref, = torch.autograd.grad(F.conv2d(x, wr, padding=k // 2), wr, gy)
    a = C._wgrad_raw(x, gy, k, 1.0)
    assert _rel(a, ref) < 2e-5
    assert torch.equal(a, C._wgrad_raw(x, gy, k, 1.0))


def test_conv_bias_act_backward_planes_vs_fp32_operands():
    """ConvBiasActFn.backward with the planes fast path on/off: same gx, gw, gb (both are bf16x3 arithmetic)."""
    from animeface_b200.ops import conv2d as C
    _setup()
    x = _cl(4, 64, 32, 32, seed=9).requires_grad_(True)
    w = torch.randn(128, 64, 3, 3, device=DEV, requires_grad=True)
    b = (torch.randn(128, device=DEV) * 0.2).requires_grad_(True)
    gy = _cl(4, 128, 32, 32, seed=10)
    outs = []
    for enabled in (True, False):
        C.planes_enabled = enabled
        try:
            y = C.conv2d_bias_act(x, w, b, 0.04, 0.2)
            outs.append(torch.autograd.grad(y, (x, w, b), gy))
        finally:
            C.planes_enabled = True
    ref_y = F.leaky_relu(F.conv2d(x, w * 0.04, b, padding=1), 0.2)
    ref = torch.autograd.grad(ref_y, (x, w, b), gy)
    for a, o, r in zip(outs[0], outs[1], ref):
        assert _rel(a, o) < 1e-5 and _rel(a, r) < 3e-5, (_rel(a, o), _rel(a, r))


def test_modulated_conv_backward_planes_vs_fp32_operands():
    from animeface_b200.ops import conv2d as C
    _setup()
    n, ci, co, hw = 4, 64, 64, 32
    x = _cl(n, ci, hw, hw, seed=11).requires_grad_(True)
    w = torch.randn(co, ci, 3, 3, device=DEV, requires_grad=True)
    s = (torch.rand(n, ci, device=DEV) + 0.5).requires_grad_(True)
    b = (torch.randn(1, co, 1, 1, device=DEV) * 0.2).requires_grad_(True)
    noise = torch.randn(n, 1, hw, hw, device=DEV)
    gy = _cl(n, co, hw, hw, seed=12)
    outs = []
    for enabled in (True, False):
        C.planes_enabled = enabled
        try:
            y = C.modulated_conv2d(x, w, s, b, noise, True, 0.2)
            outs.append(torch.autograd.grad(y, (x, w, s, b), gy))
        finally:
            C.planes_enabled = True
    for a, o in zip(*outs):
        assert _rel(a, o) < 2e-5, _rel(a, o)


@pytest.mark.parametrize('ci,co,hw', [(64, 128, 32), (32, 64, 64), (128, 128, 16)])
def test_dblock_single_node_fast_backward_vs_composed(ci, co, hw):
    """DBlockFn: pooled leaky-ReLU-gradient passes + shared x planes + accumulating skip data gradient (first-order fast path)
    against the composition of the differentiable families (planes off), and against plain torch."""
    import math
    from animeface_b200.ops import conv2d as C
    _setup()
    g = torch.Generator(device=DEV).manual_seed(ci + co)
    x = _cl(4, ci, hw, hw, seed=21).requires_grad_(True)
    mk = lambda *s: torch.randn(*s, device=DEV, generator=g).requires_grad_(True)
    w1, w2, ws = mk(co, ci, 3, 3), mk(co, co, 3, 3), mk(co, ci, 1, 1)
    b1, b2, bs = mk(co), mk(co), mk(co)
    c1, c2, cs = 1 / math.sqrt(ci * 9), 1 / math.sqrt(co * 9), 1 / math.sqrt(ci)
    gy = _cl(4, co, hw // 2, hw // 2, seed=22)
    params = (x, w1, b1, w2, b2, ws, bs)
    outs = []
    for enabled in (True, False):
        C.planes_enabled = enabled
        try:
            y = C.dblock(x, w1, b1, w2, b2, ws, bs, c1, c2, cs)
            outs.append((y.detach(),) + torch.autograd.grad(y, params, gy))
        finally:
            C.planes_enabled = True
    h = F.leaky_relu(F.conv2d(x, w1 * c1, b1, padding=1), 0.2)
    h = F.leaky_relu(F.conv2d(h, w2 * c2, b2, padding=1), 0.2)
    ref_y = (F.avg_pool2d(h, 2) + F.avg_pool2d(F.conv2d(x, ws * cs, bs), 2)) / math.sqrt(2.0)
    ref = (ref_y.detach(),) + torch.autograd.grad(ref_y, params, gy)
    # fast vs composed share the forward (same leaky-ReLU masks): they must agree to kernel round-off.  Against torch the masks
    # can differ in the sign of a handful of near-zero pre-activations (each flip moves ~600 gradient elements by a few
    # percent of their scale), so that comparison is in the L2 norm: a wrong tap / layout / scale would be O(1) there.
    l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    for name, a, o, r in zip(('y', 'gx', 'gw1', 'gb1', 'gw2', 'gb2', 'gws', 'gbs'), outs[0], outs[1], ref):
        assert _rel(a, o) < 2e-5 and l2(a, r) < 5e-3, (name, _rel(a, o), l2(a, r))
    assert _rel(outs[0][0], ref[0]) < 5e-5
