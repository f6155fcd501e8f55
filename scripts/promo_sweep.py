"""Precision / speed of the fp32-class halo convolution as a function of the promotion interval (env SG2_PROMO_TAPS,
read once by the library): error vs an fp64 reference and the time of three full-size forward layers."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import conv2d as C                                  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
dev = 'cuda'
pt = os.environ.get('SG2_PROMO_TAPS', '2')
out = [f'promo_taps={pt}']
with torch.no_grad():
    for (n, ci, co, hw) in [(4, 512, 512, 16), (4, 64, 64, 64), (2, 256, 256, 32)]:
        g = torch.Generator(device=dev).manual_seed(ci)
        x = torch.randn(n, ci, hw, hw, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
        w = torch.randn(co, ci, 3, 3, device=dev, generator=g)
        ref64 = F.conv2d(x.double(), (w * 0.05).double(), padding=1)
        y = C._conv_raw(x, w, 0.05, False, impl=5)
        y32 = F.conv2d(x, w * 0.05, padding=1)
        e = float((y.double() - ref64).abs().max() / ref64.abs().max())
        e32 = float((y32.double() - ref64).abs().max() / ref64.abs().max())
        out.append(f'{ci}->{co}@{hw}: err {e:.2e} (torch fp32 {e32:.2e})')
    for (ci, co, r) in [(64, 64, 256), (128, 128, 128), (512, 512, 32)]:
        x = torch.randn(32, ci, r, r, device=dev).contiguous(memory_format=torch.channels_last)
        w = torch.randn(co, ci, 3, 3, device=dev)
        for _ in range(3):
            C._conv_raw(x, w, 0.1, False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            C._conv_raw(x, w, 0.1, False)
        e1.record()
        torch.cuda.synchronize()
        out.append(f'fwd {ci}->{co}@{r}: {e0.elapsed_time(e1) / 5:.3f} ms')
print(' | '.join(out), flush=True)
