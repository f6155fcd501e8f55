"""R1-pattern second-order parameter gradients of the TAIL of the discriminator (last k DBlocks + minibatch-stddev + epilogue)
at B = 32: product modules vs the same arithmetic in plain torch fp32 and fp64.  Bisects where a deviation enters."""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.model import DBlock, Discriminator, MiniBatchStdDev      # noqa: E402
from animeface_b200.train import TrainConfig, build_models                    # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = 'cuda'


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def ref_tail(mods, x, dt):
    c = lambda t: t.detach().to(dt)
    params = []

    def P(t):
        p = c(t).requires_grad_(True)
        params.append(p)
        return p

    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, DBlock):
            sk = m.skip
            t = F.conv2d(x * sk.coef, P(sk.layer.weight), P(sk.layer.bias))
            for j in (0, 2):
                cv = m.block[j]
                x = F.leaky_relu(F.conv2d(x * cv.coef, P(cv.layer.weight), P(cv.layer.bias), padding=1), 0.2)
            x = (F.avg_pool2d(x, 2) + F.avg_pool2d(t, 2)) / math.sqrt(2.0)
            i += 1
        elif isinstance(m, MiniBatchStdDev):
            B, C, H, W = x.shape
            G = 4
            y = x.view(G, B // G, C, H, W)
            y = y - y.mean(0, keepdim=True)
            y = (y.square().mean(0) + 1e-4).sqrt().mean([1, 2, 3], keepdim=True).repeat(G, 1, H, W)
            x = torch.cat([x, y], 1)
            i += 1
        elif hasattr(m, 'layer') and isinstance(m.layer, torch.nn.Conv2d):
            x = F.leaky_relu(F.conv2d(x * m.coef, P(m.layer.weight), P(m.layer.bias), padding=1), 0.2)
            i += 2
        elif hasattr(m, 'layer'):
            x = F.linear(x * m.coef, P(m.layer.weight), P(m.layer.bias))
            if i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.LeakyReLU):
                x = F.leaky_relu(x, 0.2)
                i += 1
            i += 1
        else:
            x = x.reshape(x.size(0), -1)
            i += 1
    return x, params


def main():
    torch.manual_seed(3)
    cfg = TrainConfig(batch_size=32)
    _, _, D = build_models(cfg, DEV)
    mods_all = list(D.blocks)
    for k in (0, 1, 2, 3):
        mods = mods_all[6 - k:]
        ch = mods[0].block[0].layer.in_channels if k else 512
        res = 4 * 2 ** k
        x0 = torch.randn(32, ch, res, res, device=DEV)
        # product path: the same module objects run through Discriminator.forward's fusion logic
        Dp = Discriminator.__new__(Discriminator)
        torch.nn.Module.__init__(Dp)
        Dp.blocks = torch.nn.Sequential(*mods)

        def run_product(x):
            x = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
            h = x
            # reuse Discriminator.forward minus from_rgb
            import types
            src = Discriminator.forward
            class _Id(torch.nn.Module):
                pass
            out = _forward_blocks(Dp, h)
            g, = torch.autograd.grad(out.sum(), x, create_graph=True)
            pen = g.reshape(32, -1).norm(2, dim=1).pow(2).mean() / 2
            ps = [p for p in Dp.parameters()]
            return pen, torch.autograd.grad(pen, ps, allow_unused=True), [n for n, _ in Dp.named_parameters()]

        def run_ref(dt):
            x = x0.detach().to(dt).requires_grad_(True)
            out, ps = ref_tail(mods, x, dt)
            g, = torch.autograd.grad(out.sum(), x, create_graph=True)
            pen = g.reshape(32, -1).norm(2, dim=1).pow(2).mean() / 2
            return pen, torch.autograd.grad(pen, ps, allow_unused=True)

        pen_p, gp, names = run_product(x0)
        pen32, g32 = run_ref(torch.float32)
        pen64, g64 = run_ref(torch.float64)
        # ref_tail orders DBlock params skip-first; product named_parameters order is block.0, block.2, skip: match by shape multiset per module is fragile,
        # so compare as sorted-by-norm lists of the fp64 values
        print(f'tail k={k} input [{ch},{res},{res}]: pen ours {rel(pen_p, pen64):.2e} torch32 {rel(pen32, pen64):.2e}')
        key = lambda t: (tuple(t.shape), round(float(t.double().abs().sum()), 3))
        for n, a in zip(names, gp):
            if a is None:
                continue
            cands = [i for i in range(len(g64)) if g64[i] is not None and tuple(g64[i].shape) == tuple(a.shape)]
            best = min(cands, key=lambda i: rel(a, g64[i]))
            print(f'    {n:32s} ours-fp64 {rel(a, g64[best]):.2e}   torch32-fp64 {rel(g32[best], g64[best]):.2e}', flush=True)


def _forward_blocks(Dp, x):
    """Discriminator.forward without the from_rgb layer."""
    import animeface_b200.model as M
    from animeface_b200.ops import conv2d as C
    from animeface_b200.ops.bias_act import bias_act
    from animeface_b200.ops.linear import linear_bias_act
    mods = list(Dp.blocks)
    i = 0
    while i < len(mods):
        m = mods[i]
        fuse = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.LeakyReLU)
        if (isinstance(m, M.MiniBatchStdDev) and i + 2 < len(mods) and isinstance(mods[i + 1], M.ELR)
                and isinstance(mods[i + 1].layer, torch.nn.Conv2d) and isinstance(mods[i + 2], torch.nn.LeakyReLU) and x.shape[1] % 32 == 0):
            conv = mods[i + 1]
            c_in = x.shape[1]
            y = m(x)
            w = conv.layer.weight
            z = C.conv2d(x, w[:, :c_in].contiguous(), conv.coef) + C.conv2d(y[:, c_in:], w[:, c_in:].contiguous(), conv.coef)
            x = bias_act(z, conv.layer.bias, act='lrelu', alpha=M.SLOPE, gain=1.0)
            i += 3
            continue
        if isinstance(m, M.ELR) and isinstance(m.layer, torch.nn.Conv2d):
            x = C.conv2d_bias_act(x, m.layer.weight, m.layer.bias, m.coef, M.SLOPE if fuse else None)
            i += 2 if fuse else 1
        elif isinstance(m, M.ELR):
            x = linear_bias_act(x, m.layer.weight, m.layer.bias, m.coef, 1.0, M.SLOPE if fuse else None)
            i += 2 if fuse else 1
        else:
            x = m(x)
            i += 1
    return x


if __name__ == '__main__':
    main()
