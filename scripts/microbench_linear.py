"""Times the fully connected kernels (csrc/linear.cu) on the shapes of the path: mapping 512->512, style affines 512->Ci,
discriminator epilogue 8192->512 and 512->1 (CUDA events, median of 20, warm L2 like inside the step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import linear as L                                 # noqa: E402

DEV = 'cuda'


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    for B, K, N in [(32, 512, 512), (32, 512, 64), (32, 512, 32), (64, 8192, 512), (64, 512, 1), (32, 8192, 512)]:
        x, w, b = torch.randn(B, K, device=DEV), torch.randn(N, K, device=DEV), torch.randn(N, device=DEV)
        y = L._fwd_raw(x, w, b, 0.1, 1.4, 0.2)
        gy = torch.randn(B, N, device=DEV)
        t_f = timeit(lambda: L._fwd_raw(x, w, b, 0.1, 1.4, 0.2))
        t_dx = timeit(lambda: L._dx_raw(gy, y, w, 0.1, 1.4, 0.2))
        t_dw = timeit(lambda: L._dw_raw(gy, y, x, 0.1, 1.4, 0.2, True))
        wbytes = N * K * 4
        print(f'B={B:3d} K={K:5d} N={N:4d}  fwd {t_f:7.1f} us  dx {t_dx:7.1f} us  dw {t_dw:7.1f} us   (W = {wbytes / 1e6:.2f} MB: {wbytes / 6.5e6:.1f} us at HBM rate)')


if __name__ == '__main__':
    with torch.no_grad():
        main()
