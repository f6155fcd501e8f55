"""GPU parity of the fused DiffAugment kernels (csrc/diffaug.cu) against the per-op restatement of
thirdparty/diffaugment/DiffAugment.py (same draws, replayed): forward, gradient, and the gradient of the gradient."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _draws(policy, B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for p in policy.split(','):
        if p == 'color':
            out += [torch.rand(B, 1, 1, 1, generator=g) for _ in range(3)]
        elif p == 'translation':
            sh, sw = int(H * 0.125 + 0.5), int(W * 0.125 + 0.5)
            out += [torch.randint(-sh, sh + 1, (B, 1, 1), generator=g), torch.randint(-sw, sw + 1, (B, 1, 1), generator=g)]
        else:
            ch, cw = int(H * 0.5 + 0.5), int(W * 0.5 + 0.5)
            out += [torch.randint(0, H + (1 - ch % 2), (B, 1, 1), generator=g), torch.randint(0, W + (1 - cw % 2), (B, 1, 1), generator=g)]
    return out


@pytest.mark.parametrize('policy', ['color,translation', 'color', 'translation', 'color,translation,cutout', 'translation,cutout'])
@pytest.mark.parametrize('shape', [(8, 3, 32, 32), (5, 3, 24, 40), (4, 1, 16, 16)])
def test_fused_diffaugment_matches_per_op_path(policy, shape):
    from animeface_b200 import diffaugment as DA
    from animeface_b200 import rng
    B, C, H, W = shape
    x = torch.randn(*shape, device=DEV).requires_grad_(True)
    gy = torch.randn(*shape, device=DEV)
    v = torch.randn(*shape, device=DEV)
    res = []
    for fused in (True, False):
        with rng.replay(_draws(policy, B, H, W, 7)) as q:
            if fused:
                assert DA._fused_ok(x, policy.split(','))
                y = DA.DiffAugment(x, policy)
            else:
                y = x
                for p in policy.split(','):
                    for f in DA.AUGMENT_FNS[p]:
                        y = f(y)
            assert q.remaining == 0
        gyr = gy.clone().requires_grad_(True)
        gx, = torch.autograd.grad(y, x, gyr, create_graph=True)
        ggy, = torch.autograd.grad(gx, gyr, v)                 # d(A^T gy . v)/d gy = A v
        res.append((y.detach(), gx.detach(), ggy))
    for name, a, r in zip(('y', 'gx', 'A v'), *res):
        assert _rel(a, r) < 2e-6, (name, _rel(a, r))


def test_fused_diffaugment_full_size_is_deterministic():
    from animeface_b200 import diffaugment as DA
    from animeface_b200 import rng
    x = torch.rand(32, 3, 256, 256, device=DEV) * 2 - 1
    outs = []
    for _ in range(2):
        with rng.replay(_draws('color,translation', 32, 256, 256, 3)):
            outs.append(DA.DiffAugment(x, 'color,translation'))
    assert torch.equal(outs[0], outs[1]) and torch.isfinite(outs[0]).all()
