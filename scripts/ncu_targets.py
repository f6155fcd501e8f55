"""The launches `ncu --set full` is pointed at (one warm-up + one profiled launch per kernel, full-size layer shapes).
Usage (GPU box):
    ncu --set full --import-source on --clock-control none -k regex:'conv_halo_kernel|conv_wgrad_tc|up2x_adj' \
        -o gpurun_out/x python scripts/ncu_targets.py [names...]
Each target launches its kernel twice (ncu profiles both; read the second)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import conv2d as C                                  # noqa: E402
from animeface_b200.ops.resample import Up2xAdjFn                           # noqa: E402

DEV, B = 'cuda', 32
SHAPES = {'64': (64, 64, 256), '128': (128, 128, 128), '256': (256, 256, 64), '512': (512, 512, 32), '32': (32, 64, 256)}


def cl(*shape):
    return torch.randn(*shape, device=DEV).contiguous(memory_format=torch.channels_last)


def main():
    which = sys.argv[1:] or ['fwd64', 'fwd128', 'dgrad64', 'wgrad64', 'wgrad256', 'up2x_adj']
    with torch.no_grad():
        for name in which:
            if name == 'up2x_adj':
                g = cl(B, 64, 256, 256)
                for _ in range(2):
                    Up2xAdjFn.apply(g, True)
            else:
                kind = name.rstrip('0123456789')
                ci, co, r = SHAPES[name[len(kind):]]
                w = torch.randn(co, ci, 3, 3, device=DEV)
                if kind == 'fwd':
                    x = cl(B, ci, r, r)
                    for _ in range(2):
                        C._conv_raw(x, w, 0.1, False)
                elif kind == 'dgrad':
                    x = cl(B, co, r, r)
                    for _ in range(2):
                        C._conv_raw(x, w, 0.1, True)
                elif kind == 'wgrad':
                    x, gy = cl(B, ci, r, r), cl(B, co, r, r)
                    for _ in range(2):
                        C._wgrad_raw(x, gy, 3, 0.1)
                else:
                    raise SystemExit(f'unknown target {name}')
            torch.cuda.synchronize()
            print(name, 'done', flush=True)


if __name__ == '__main__':
    main()
