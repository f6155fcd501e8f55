"""The demodulation coefficient kernels (csrc/demod.cu) against the tensor expression they replace
(implementations/StyleGAN2/model.py:115-120 reduced to the [B,Co] coefficient): value, d/dw, d/ds; repeat runs bit-identical."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize('shape', [(32, 512, 512, 3), (8, 32, 64, 3), (5, 3, 32, 1), (64, 96, 203, 1)])
def test_demod_matches_the_tensor_expression(shape):
    from animeface_b200.ops.conv2d import DemodFn
    B, co, ci, k = shape
    g = torch.Generator().manual_seed(co + ci)
    w0 = torch.randn(co, ci, k, k, generator=g).to(DEV)
    s0 = (torch.randn(B, ci, generator=g) + 1).to(DEV)
    gd = torch.randn(B, co, generator=g).to(DEV)
    coef, eps = 1.0 / (ci * k * k) ** 0.5, 1e-4
    w, s = w0.clone().requires_grad_(True), s0.clone().requires_grad_(True)
    d = DemodFn.apply(w, s, coef, eps)
    gw, gs = torch.autograd.grad(d, (w, s), gd)
    wr, sr = w0.double().requires_grad_(True), s0.double().requires_grad_(True)
    dr = torch.rsqrt(torch.matmul(sr.square(), wr.square().sum((2, 3)).t()) * (coef * coef) + eps)
    gwr, gsr = torch.autograd.grad(dr, (wr, sr), gd.double())
    assert _rel(d.double(), dr) < 2e-6
    assert _rel(gw.double(), gwr) < 1e-5 and _rel(gs.double(), gsr) < 1e-5
    d2 = DemodFn.apply(w, s, coef, eps)
    gw2, gs2 = torch.autograd.grad(d2, (w, s), gd)
    assert torch.equal(d, d2) and torch.equal(gw, gw2) and torch.equal(gs, gs2)
