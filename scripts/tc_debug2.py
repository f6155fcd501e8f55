import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import conv2d as C
torch.backends.cudnn.allow_tf32 = False
dev = 'cuda'
for (n, ci, co, k, hw) in [(8, 32, 32, 3, 16), (8, 32, 32, 3, 8), (8, 16, 32, 3, 16)]:
    g = torch.Generator(device=dev).manual_seed(1)
    w = torch.randn(co, ci, k, k, device=dev, generator=g)
    gy = torch.randn(n, co, hw, hw, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    x = torch.randn(n, ci, hw, hw, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    coef = 0.06
    for sparse in (False, True):
        gyy = gy * (torch.rand_like(gy) > 0.5) if sparse else gy
        ref = F.conv_transpose2d(gyy, w * coef, padding=1)
        for impl in (1, 2):
            try:
                out = C._conv_raw(gyy, w, coef, True, impl=impl)
            except RuntimeError as e:
                print('skip', impl, str(e)[:60]); continue
            d = (out - ref).abs()
            print(f'n={n} {ci}->{co}@{hw} sparse={sparse} impl={impl}: dgrad err {float(d.max() / ref.abs().max()):.2e}  argmax {tuple(int(v) for v in torch.unravel_index(d.argmax(), d.shape))}')
        b = torch.randn(co, device=dev, generator=g)
        ref = F.leaky_relu(F.conv2d(x, w * coef, padding=1) + b[None, :, None, None], 0.2)
        for impl in (1, 2):
            try:
                out = C._conv_raw(x, w, coef, False, bias=b, slope=0.2, impl=impl)
            except RuntimeError as e:
                continue
            d = (out - ref).abs()
            print(f'      fwd impl={impl} err {float(d.max() / ref.abs().max()):.2e} flips {int(((out > 0) != (ref > 0)).sum())}')
