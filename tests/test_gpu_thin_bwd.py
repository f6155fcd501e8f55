"""The fused backward of a thin-input 1x1 convolution + bias + leaky ReLU (csrc/thin_bwd.cu: Discriminator.from_rgb,
implementations/StyleGAN2/model.py:383-384) against torch autograd in fp64: gx, gw, gb; ragged pixel counts; repeat runs identical."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize('shape', [(4, 3, 32, 64, 64), (2, 3, 32, 37, 29), (3, 1, 8, 16, 16), (2, 4, 64, 20, 20)])
@pytest.mark.parametrize('gain', [1.0, 1.4142])
def test_thin_in_backward_matches_autograd(shape, gain):
    from animeface_b200.ops.conv2d import conv2d_bias_act
    n, ci, co, h, w_ = shape
    g = torch.Generator().manual_seed(ci * 100 + co)
    x0 = torch.randn(n, ci, h, w_, generator=g).to(DEV)
    w0 = torch.randn(co, ci, 1, 1, generator=g).to(DEV)
    b0 = torch.randn(co, generator=g).to(DEV)
    gy = torch.randn(n, co, h, w_, generator=g).to(DEV)
    coef = 1.0 / ci ** 0.5
    x, w, b = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    y = conv2d_bias_act(x, w, b, coef, 0.2, gain)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    xr, wr, br = x0.double().requires_grad_(True), w0.double().requires_grad_(True), b0.double().requires_grad_(True)
    yr = F.leaky_relu(F.conv2d(xr, wr * coef) + br.reshape(1, -1, 1, 1), 0.2) * gain
    gxr, gwr, gbr = torch.autograd.grad(yr, (xr, wr, br), gy.double())
    assert _rel(gx, gxr) < 1e-5 and _rel(gw, gwr) < 1e-5 and _rel(gb, gbr) < 1e-5, (_rel(gx, gxr), _rel(gw, gwr), _rel(gb, gbr))
    y2 = conv2d_bias_act(x, w, b, coef, 0.2, gain)
    gx2, gw2, gb2 = torch.autograd.grad(y2, (x, w, b), gy)
    assert torch.equal(gx, gx2) and torch.equal(gw, gw2) and torch.equal(gb, gb2)
