"""tcgen05 conv kernels vs torch fp32 conv (TF32 off): prints max relative-to-scale error per case."""
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


cases = [(2, 64, 64, 3, 16), (8, 32, 64, 3, 16), (8, 64, 64, 3, 32), (3, 64, 128, 3, 8), (8, 128, 128, 1, 16), (2, 32, 32, 3, 64),
         (8, 512, 512, 3, 4), (2, 64, 32, 3, 128), (1, 256, 256, 3, 16), (4, 64, 64, 1, 256), (5, 96, 160, 3, 8)]
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
for idx, (n, ci, co, k, hw) in enumerate(cases):
    g = torch.Generator(device=dev).manual_seed(idx)
    x = torch.randn(n, ci, hw, hw, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn(co, ci, k, k, device=dev, generator=g)
    s = torch.randn(n, ci, device=dev, generator=g)
    d = torch.rand(n, co, device=dev, generator=g) + 0.5
    b = torch.randn(co, device=dev, generator=g)
    nz = torch.randn(n, 1, hw, hw, device=dev, generator=g)
    gy = torch.randn(n, co, hw, hw, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    coef = 0.05
    out = [f'case {idx} n={n} {ci}->{co} k{k} @{hw}:']
    fwd_ok = co in (32, 64) or co % 128 == 0
    if which in ('all', 'fwd') and fwd_ok:
        ref0 = F.conv2d(x, w * coef, padding=k // 2)
        ref1 = F.leaky_relu(F.conv2d(x * s[:, :, None, None], w * coef, padding=k // 2) * d[:, :, None, None] + b[None, :, None, None] + nz, 0.2)
        ref64 = F.leaky_relu(F.conv2d((x * s[:, :, None, None]).double(), (w * coef).double(), padding=k // 2) * d[:, :, None, None].double() + b[None, :, None, None].double() + nz.double(), 0.2)
        e32 = float((ref1.double() - ref64).abs().max() / ref64.abs().max())
        for impl in (4, 5):
            y0 = C._conv_raw(x, w, coef, False, impl=impl)
            y1 = C._conv_raw(x, w, coef, False, in_scale=s, out_scale=d, bias=b, noise=nz, slope=0.2, impl=impl)
            e64 = float((y1.double() - ref64).abs().max() / ref64.abs().max())
            out.append(f'impl{impl} fwd {err(y0, ref0):.1e} fused-vs-fp64 {e64:.1e} flips {int(((y1 > 0) != (ref64 > 0)).sum())}')
        out.append(f'(torch fp32 vs fp64 {e32:.1e})')
        if ci in (32, 64) or ci % 128 == 0:
            ref2 = F.conv_transpose2d(gy, w * coef, padding=k // 2)
            out.append(f'dgrad {err(C._conv_raw(gy, w, coef, True, impl=4), ref2):.1e}')
            if hw % 16 == 0:
                out.append(f'dgrad-halo {err(C._conv_raw(gy, w, coef, True, impl=4), ref2):.1e}')
    if which in ('all', 'wgrad'):
        xr = x.detach().clone().requires_grad_(False)
        wr = w.detach().clone().requires_grad_(True)
        yr = F.conv2d(xr, wr * coef, padding=k // 2)
        ref_w, = torch.autograd.grad(yr, wr, gy)
        dw = C._wgrad_raw(x, gy, k, coef, impl=4)
        yr2 = F.conv2d(xr * s[:, :, None, None], wr * coef, padding=k // 2) * d[:, :, None, None]
        ref_w2, = torch.autograd.grad(yr2, wr, gy)
        dw2 = C._wgrad_raw(x, gy, k, coef, in_scale=s, out_scale=d, impl=4)
        out.append(f'wgrad {err(dw, ref_w):.1e} scaled {err(dw2, ref_w2):.1e}')
    torch.cuda.synchronize()
    print('  '.join(out), flush=True)
