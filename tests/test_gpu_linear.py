"""GPU parity of the fully connected kernels (csrc/linear.cu, ops/linear.py) against torch fp32 (TF32 off): forward, first-order
gradients, the R1-style second order through the closed family, PixelNorm, and the Mapping network built on them."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _ref(x, w, b, coef, gain, slope):
    t = (F.linear(x * coef, w, b)) * gain
    return F.leaky_relu(t, slope) if slope is not None else t


CASES = [(32, 512, 512, 0.01, 0.2), (64, 8192, 512, 1.0, 0.2), (64, 512, 1, 1.0, None), (5, 70, 33, 0.5, 0.2), (3, 64, 48, 1.0, None),
         (17, 516, 130, 2.0, 0.2)]


@pytest.mark.parametrize('B,K,N,gain,slope', CASES)
def test_linear_forward_backward_and_second_order(B, K, N, gain, slope):
    from animeface_b200.ops.linear import linear_bias_act
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(B * 7 + N)
    x = torch.randn(B, K, device=DEV, generator=g).requires_grad_(True)
    w = torch.randn(N, K, device=DEV, generator=g).requires_grad_(True)
    b = (torch.randn(N, device=DEV, generator=g) * 0.3).requires_grad_(True)
    gy = torch.randn(B, N, device=DEV, generator=g)
    coef = 1.0 / K ** 0.5
    outs = []
    for fn in (linear_bias_act, _ref):
        y = fn(x, w, b, coef, gain, slope)
        gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy, create_graph=True)
        pen = gx.square().sum()
        g2 = torch.autograd.grad(pen, (w, b), allow_unused=True)
        outs.append((y, gx, gw, gb, g2[0]))
    for name, a, r in zip(('y', 'gx', 'gw', 'gb', 'r1-style d/dw'), outs[0], outs[1]):
        assert _rel(a, r) < 2e-5, (name, _rel(a, r))
    # fused first-order backward (no graph) == the composed one
    y = linear_bias_act(x, w, b, coef, gain, slope)
    fx, fw, fb = torch.autograd.grad(y, (x, w, b), gy)
    for a, r in zip((fx, fw, fb), outs[1][1:4]):
        assert _rel(a, r) < 2e-5
    assert torch.equal(fw, torch.autograd.grad(linear_bias_act(x, w, b, coef, gain, slope), w, gy)[0])      # deterministic


def test_pixelnorm_and_mapping_network():
    from animeface_b200.model import Mapping, init_weight_N01
    from animeface_b200.ops.linear import pixel_norm
    import functools
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    z = torch.randn(32, 512, device=DEV)
    assert _rel(pixel_norm(z), z / (z.pow(2).mean(dim=1, keepdim=True).sqrt() + 1e-4)) < 1e-6
    m = Mapping(512, 8, True, 0.01).to(DEV)
    m.apply(functools.partial(init_weight_N01, lr=0.01))
    for p in m.parameters():
        if p.ndim == 1:
            p.data.normal_(0, 0.5)
    y = m(z)
    # the reference's module arithmetic (model.py:29-37, 71-78, 253-282) in plain torch
    x = z / (z.pow(2).mean(dim=1, keepdim=True).sqrt() + 1e-4)
    for i in range(0, 16, 2):
        lin = m.map[i].linear
        x = F.leaky_relu(F.linear(x * lin.coef, lin.layer.weight, lin.layer.bias) * 0.01, 0.2)
    assert _rel(y, x) < 1e-5
    gy = torch.randn_like(y)
    ps = list(m.parameters())
    ga = torch.autograd.grad(y, ps, gy)
    gb = torch.autograd.grad(x, ps, gy)
    for (n, _), a, r in zip(m.named_parameters(), ga, gb):
        assert _rel(a, r) < 2e-5, (n, _rel(a, r))
