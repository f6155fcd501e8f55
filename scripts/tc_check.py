"""tcgen05 conv kernel vs torch fp32 conv (TF32 off): prints max relative-to-scale error per case."""
import os
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import conv2d as C

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda'


def err(a, b):
    return float((a - b).abs().max() / b.abs().max())


cases = [(8, 32, 64, 3, 16), (8, 64, 64, 3, 32), (3, 64, 128, 3, 8), (8, 128, 128, 1, 16), (2, 32, 32, 3, 64),
         (8, 512, 512, 3, 4), (2, 64, 32, 3, 128), (1, 256, 256, 3, 16), (4, 64, 64, 1, 256)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for idx, (n, ci, co, k, hw) in enumerate(cases):
    if only is not None and idx != only:
        continue
    g = torch.Generator(device=dev).manual_seed(idx)
    x = torch.randn(n, ci, hw, hw, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn(co, ci, k, k, device=dev, generator=g)
    s = torch.randn(n, ci, device=dev, generator=g)
    d = torch.rand(n, co, device=dev, generator=g) + 0.5
    b = torch.randn(co, device=dev, generator=g)
    nz = torch.randn(n, 1, hw, hw, device=dev, generator=g)
    coef = 0.05
    ref0 = F.conv2d(x, w * coef, padding=k // 2)
    y0 = C._conv_raw(x, w, coef, False, impl=2)
    torch.cuda.synchronize()
    e0 = err(y0, ref0)
    ref1 = F.leaky_relu(F.conv2d(x * s[:, :, None, None], w * coef, padding=k // 2) * d[:, :, None, None] + b[None, :, None, None] + nz, 0.2)
    y1 = C._conv_raw(x, w, coef, False, in_scale=s, out_scale=d, bias=b, noise=nz, slope=0.2, impl=2)
    e1 = err(y1, ref1)
    gy = torch.randn(n, co, hw, hw, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    ref2 = F.conv_transpose2d(gy, w * coef, padding=k // 2)
    y2 = C._conv_raw(gy, w, coef, True, impl=2)
    e2 = err(y2, ref2)
    ysimt = C._conv_raw(x, w, coef, False, impl=1)
    print(f'case {idx} n={n} {ci}->{co} k{k} @{hw}: fwd {e0:.2e}  fused {e1:.2e}  dgrad {e2:.2e}  (simt fwd {err(ysimt, ref0):.2e})', flush=True)
