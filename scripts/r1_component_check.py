"""Second-order (R1 pattern) error of single components at the shapes of BASELINE config 2, B = 32: ours vs torch fp64,
next to torch fp32 vs fp64.  Pattern: y = f(x); gx = d sum(y * gy)/dx (create_graph); pen = sum(gx^2); d pen / d params."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops.mbstd import minibatch_stddev        # noqa: E402
from oracle import sg2_torch as T                             # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = 'cuda'


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def ref_mbstd(x, G=4, eps=1e-4):
    B, C, H, W = x.shape
    M = B // G
    y = x.view(G, M, C, H, W)
    y = y - y.mean(0, keepdim=True)
    y = (y.square().mean(0) + eps).sqrt().mean([1, 2, 3], keepdim=True)       # [M,1,1,1]
    y = y.repeat(G, 1, H, W)
    return torch.cat([x, y], 1)


def second_order(fn, x, gy, extra):
    x = x.detach().requires_grad_(True)
    extra = [e.detach().requires_grad_(True) for e in extra]
    y = fn(x, *extra)
    gx, = torch.autograd.grad((y * gy.to(y.dtype)).sum(), x, create_graph=True)
    pen = gx.square().sum()
    return [gx.detach()] + [g.detach() for g in torch.autograd.grad(pen, [x] + extra, allow_unused=True) if g is not None]


def main():
    torch.manual_seed(0)
    for B in (4, 8, 32):
        x = torch.randn(B, 512, 4, 4, device=DEV)
        bias = torch.randn(1, 512, 1, 1, device=DEV) * 0.1
        gy = torch.randn(B, 513, 4, 4, device=DEV)
        ours = second_order(lambda t, b: minibatch_stddev(F.leaky_relu(t + b, 0.2), 4), x, gy, [bias])
        r32 = second_order(lambda t, b: ref_mbstd(F.leaky_relu(t + b, 0.2)), x, gy, [bias])
        r64 = second_order(lambda t, b: ref_mbstd(F.leaky_relu(t + b, 0.2)), x.double(), gy.double(), [bias.double()])
        print(f'mbstd B={B}: ' + '  '.join(f'{n}: ours {rel(a, c):.2e} torch32 {rel(b, c):.2e}' for n, a, b, c in zip(('gx', 'd_x', 'd_bias'), ours, r32, r64)), flush=True)


if __name__ == '__main__':
    main()
