"""Times the backward-prologue passes (planes.cu) on the shapes of the first discriminator block at the D-phase batch (64)
and the generator's 256^2 layer (32): CUDA events, L2 flushed, median of 10; GB/s from the algorithmic bytes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import conv2d as C                                  # noqa: E402

DEV = 'cuda'
flush = None


def timeit(fn, iters=10):
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


def cl(*s):
    return torch.randn(*s, device=DEV).contiguous(memory_format=torch.channels_last)


def main():
    for n, c, r in [(64, 64, 256), (64, 32, 256), (64, 128, 128), (32, 32, 256)]:
        y, gy, gp = cl(n, c, r, r), cl(n, c, r, r), cl(n, c, r // 2, r // 2)
        e = n * c * r * r
        d, b, nz = torch.rand(n, c, device=DEV) + 0.5, torch.randn(c, device=DEV), torch.randn(n, 1, r, r, device=DEV)
        cases = [('plain gy,y -> planes', lambda: C._bwd_prep_planes(gy, y, 0.2), 12 * e),
                 ('pooled gy,y -> planes', lambda: C._bwd_prep_planes(gp, y, 0.2, pooled=True, gscale=0.2), 9 * e),
                 ('pooled gy -> planes', lambda: C._bwd_prep_planes(gp, None, None, pooled=True, gscale=0.2), 5 * e),
                 ('modconv gy,y,noise,d -> planes', lambda: C._bwd_prep_planes(gy, y, 0.2, noise=nz, bias=b, d=d), 12 * e),
                 ('split x -> planes', lambda: C._split_planes(y), 8 * e)]
        for name, fn, nbytes in cases:
            ms = timeit(fn)
            print(f'[{n},{c},{r},{r}] {name:34s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s')


if __name__ == '__main__':
    with torch.no_grad():
        main()
