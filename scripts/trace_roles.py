"""Per-role wait-cycle breakdown of the persistent tcgen05 kernels (sg2_debug_trace).
    python scripts/trace_roles.py [fwd64 dgrad64 fwd128 wgrad64 ...]
For each target: runs it once with tracing on and prints the mean over CTAs of every slot, as cycles and as a share of the
kernel's total cycles.  Slots -- halo conv: 0 patch-producer wait(slot free), 1 transform wait(patch landed),
2 transform wait(plane free), 3 transform total, 4 mma wait(plane ready), 5 mma wait(weights landed), 6 mma wait(acc free),
7 mma total, 8 epilogue wait(acc ready), 9 epilogue total, 10 weight-producer wait(stage free), 11 kernel total, 12 tiles,
16 / 17 mma warp: cycles issuing tcgen05.mma / tcgen05.commit,
13 epilogue promotion (TMEM -> registers, waits excluded), 14 epilogue finish + store, 15 epilogue staging hand-over (TMA-store
convs).  `mod<L>` = the modulated forward (style, demodulation, bias, noise, leaky-ReLU).
wgrad: 0 producer wait(stage free), 1 transform wait(boxes landed), 2 transform wait(planes free), 3 transform total,
4 mma wait(planes ready), 7 mma total, 11 kernel total, 12 chunks."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200 import _lib                                             # noqa: E402
from animeface_b200.ops import conv2d as C                                  # noqa: E402

DEV, B = 'cuda', 32
SHAPES = {'64': (64, 64, 256), '128': (128, 128, 128), '256': (256, 256, 64), '512': (512, 512, 32), '32': (32, 64, 256),
          '32x32': (32, 32, 256), '64x32': (64, 32, 256), '64x128': (64, 128, 128)}
NAMES = {0: 'prod wait slot-free', 1: 'xform wait data', 2: 'xform wait plane-free', 3: 'xform total', 4: 'mma wait planes',
         5: 'mma wait weights', 6: 'mma wait acc-free', 7: 'mma total', 8: 'epi wait acc', 9: 'epi total',
         10: 'wprod wait stage-free', 11: 'kernel total', 12: 'units', 13: 'epi promotion', 14: 'epi finish+store', 15: 'epi staging', 16: 'mma issue', 17: 'mma commit'}


def cl(*shape):
    return torch.randn(*shape, device=DEV).contiguous(memory_format=torch.channels_last)


def run(name):
    k = 3
    if name.endswith('k1'):
        name, k = name[:-2], 1
    kind = name.rstrip('0123456789x')
    ci, co, r = SHAPES[name[len(kind):]]
    w = torch.randn(co, ci, k, k, device=DEV)
    if kind == 'fwd':
        x = cl(B, ci, r, r); fn = lambda: C._conv_raw(x, w, 0.1, False)
    elif kind == 'mod':
        x = cl(B, ci, r, r)
        s, d = torch.rand(B, ci, device=DEV) + 0.5, torch.rand(B, co, device=DEV) + 0.5
        bias, nz = torch.randn(co, device=DEV), torch.randn(B, 1, r, r, device=DEV)
        fn = lambda: C._conv_raw(x, w, 0.1, False, in_scale=s, out_scale=d, bias=bias, noise=nz, slope=0.2)
    elif kind == 'dgrad':
        x = cl(B, co, r, r); fn = lambda: C._conv_raw(x, w, 0.1, True)
    else:
        x, gy = cl(B, ci, r, r), cl(B, co, r, r); fn = lambda: C._wgrad_raw(x, gy, k, 0.1)
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device=DEV)
    lib = _lib.load()
    lib.sg2_debug_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.sg2_debug_trace(None)
    t = buf.view(-1, 16 if kind == 'wgrad' else 32).cpu().double()
    t = t[t[:, 11] > 0]
    m = t.mean(0)
    tot = float(m[11])
    print(f'== {name}: {ci}->{co}@{r}  {e0.elapsed_time(e1):.3f} ms untraced, {t.shape[0]} CTAs, {tot:.0f} cycles/CTA, '
          f'{float(m[12]):.1f} units/CTA, {tot / max(float(m[12]), 1):.0f} cycles/unit')
    for i in (0, 1, 2, 3, 4, 5, 6, 7, 16, 17, 8, 9, 13, 14, 15, 10, 11):
        if float(m[i]) > 0:
            print(f'   [{i:2d}] {NAMES[i]:24s} {float(m[i]):12.0f} cyc  {100 * float(m[i]) / tot:5.1f}%')


if __name__ == '__main__':
    with torch.no_grad():
        for n in (sys.argv[1:] or ['fwd64', 'dgrad64', 'fwd128', 'dgrad128', 'fwd512', 'wgrad64', 'wgrad256', 'wgrad32']):
            run(n)
