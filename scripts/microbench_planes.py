"""Micro-benchmark of the bf16 pair-planes gradient kernels against the fp32-operand kernels they replace on the first-order
backward pass (CUDA events, L2 flushed between iterations, median of 10).  One JSON line per (layer, kernel).
Layers = the discriminator convolutions of BASELINE config 2 at the D-phase batch (real + fake = 64) and the generator's."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import conv2d as C                      # noqa: E402

DEV = 'cuda'
_flush = None


def timeit(fn, iters=10, warmup=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def cl(*shape):
    return torch.randn(*shape, device=DEV).contiguous(memory_format=torch.channels_last)


LAYERS = [  # n, ci, co, k, hw
    (64, 64, 64, 3, 256), (64, 64, 128, 3, 128), (64, 128, 128, 3, 128), (64, 128, 256, 3, 64), (64, 256, 256, 3, 64),
    (64, 256, 512, 3, 32), (64, 512, 512, 3, 32), (64, 512, 512, 3, 16), (64, 512, 512, 3, 8), (64, 64, 128, 1, 128),
    (32, 64, 64, 3, 256), (32, 128, 128, 3, 128), (32, 256, 256, 3, 64), (32, 512, 512, 3, 32),
]


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ''
    with torch.no_grad():
        for (n, ci, co, k, hw) in LAYERS:
            x, gy = cl(n, ci, hw, hw), cl(n, co, hw, hw)
            w = torch.randn(co, ci, k, k, device=DEV)
            flops = 2.0 * n * hw * hw * ci * co * k * k
            name = f'{ci}->{co} k{k} @{hw} n{n}'
            rows = {}
            if only in ('', 'wgrad'):
                rows['wgrad_fp32_operands'] = timeit(lambda: C._wgrad_raw(x, gy, k, 0.05))
                if C._planes_ok(n, hw, hw, ci, co, k, True):
                    xp, gyp = C._split_planes(x), C._split_planes(gy)
                    rows['wgrad_planes'] = timeit(lambda: C._wgrad_planes(xp, gyp, k, 0.05))
            if only in ('', 'dgrad'):
                rows['dgrad_fp32_operands'] = timeit(lambda: C._conv_raw(gy, w, 0.05, True))
                if C._planes_ok(n, hw, hw, co, ci, k, False):
                    gyp = C._split_planes(gy)
                    rows['dgrad_planes'] = timeit(lambda: C._conv_planes(gyp, w, 0.05, True))
            if only in ('', 'aux'):
                rows['split_planes(x)'] = timeit(lambda: C._split_planes(x))
                y = cl(n, co, hw, hw)
                rows['bwd_prep_planes(gy,y)'] = timeit(lambda: C._bwd_prep_planes(gy, y, 0.2))
                rows['fwd_precise'] = timeit(lambda: C._conv_raw(x, w, 0.05, False))
            for kname, ms in rows.items():
                r = dict(layer=name, kernel=kname, ms=round(ms, 4))
                if 'grad' in kname or 'fwd' in kname:
                    r['TFLOPs'] = round(flops / ms / 1e9, 1)
                print(json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
