"""filtered_lrelu: the fused kernel against the composed bias_act / upfirdn2d execution on StyleGAN3-sized layers
(B = 16, up = down = 2, 12-tap Kaiser-like filters; the radial down filter is a full 12 x 12).  CUDA events, L2 flushed, median
of 7.  GB/s counts the algorithmic bytes of the fused form: x + y (+ 1 byte per up-sampled element for the sign mask when a
backward follows; backward: dy + mask + dx)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import animeface_b200.ops.filtered_lrelu as F                               # noqa: E402

DEV = 'cuda'
flush = None


def timeit(fn, iters=7):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def filt(n, two_d):
    f = torch.hamming_window(n, periodic=False) + 0.05
    f = f / f.sum()
    if two_d:
        f = torch.outer(f, f)
    return f.to(DEV)


def main():
    for n, c, hw, two_d in [(16, 512, 52, True), (16, 256, 84, True), (16, 128, 148, True), (16, 256, 84, False), (16, 128, 148, False)]:
        fu, fd = filt(12, False), filt(12, two_d)
        pad = [9, 10, 9, 10]
        x = torch.randn(n, c, hw, hw, device=DEV)
        b = torch.randn(c, device=DEV)
        row = {}
        for fused in (True, False):
            F.fused_enabled = fused
            with torch.no_grad():
                y = F.filtered_lrelu(x, fu, fd, b, 2, 2, pad)
                row[('fwd', fused)] = timeit(lambda: F.filtered_lrelu(x, fu, fd, b, 2, 2, pad))
            xg = x.clone().requires_grad_(True)
            gy = torch.randn_like(y)

            def step():
                yy = F.filtered_lrelu(xg, fu, fd, b, 2, 2, pad)
                torch.autograd.grad(yy, xg, gy)
            row[('fwd+bwd', fused)] = timeit(step)
        F.fused_enabled = True
        z = (hw * 2 + 19 - 11) ** 2 * n * c
        fb = (x.numel() + y.numel()) * 4
        fbb = fb + z + (y.numel() + x.numel()) * 4 + z
        print(f'[{n},{c},{hw},{hw}] -> {tuple(y.shape[2:])} fd {"12x12" if two_d else "12 separable"}: '
              f'fwd fused {row[("fwd", True)]:.3f} ms ({fb / row[("fwd", True)] / 1e6:.0f} GB/s) composed {row[("fwd", False)]:.3f} ms; '
              f'fwd+bwd fused {row[("fwd+bwd", True)]:.3f} ms ({fbb / row[("fwd+bwd", True)] / 1e6:.0f} GB/s) composed {row[("fwd+bwd", False)]:.3f} ms')


if __name__ == '__main__':
    main()
