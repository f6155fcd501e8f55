"""The launches `ncu --set full` is pointed at (one warm-up + one profiled launch per kernel, full-size layer shapes, B = 32).
Usage (GPU box):
    ncu --set full --import-source on --clock-control none -k regex:'<kernels>' -o gpurun_out/x python scripts/ncu_targets.py [names...]
Each target launches its kernel twice (ncu profiles both; read the second).  Targets:
  fwd<L> / dgrad<L> / wgrad<L>     fp32-operand kernels (conv_halo.cu fp32-class / bf16x3, wgrad_tc.cu), L in SHAPES
  mod<L>                           the MODULATED forward: style scale on the activation tile, demodulation + bias + noise +
                                   leaky-ReLU in the epilogue (what north_star calls modulated_conv2d)
  dgradpl<L> / wgradpl<L>          bf16 pair-planes kernels of the first-order backward (conv_halo_pl.cu, wgrad_pl.cu)
  prep64 / preppool64 / split64    planes producers (planes.cu) on [32,64,256,256] (preppool: the pooling adjoint folded in)
  fwdk1                            1x1 forward conv 32->64 @256^2 (staged TMA-store epilogue)
  up2x_fwd / up2x_adj / avgpool    StyleGAN2 resampling (resample_sg2.cu) at U1 / U3
  u4_nhwc / u4_nchw / u4_down2     upfirdn2d register-ring kernel (upfirdn2d.cu) at U4
  bias_act / mbstd / diffaug       bias_act_vec4 on [32,64,256,256], minibatch-stddev on [32,512,4,4], DiffAugment on [32,3,256,256]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import conv2d as C                                  # noqa: E402
from animeface_b200.ops import upfirdn2d as U                               # noqa: E402
from animeface_b200.ops.bias_act import bias_act                            # noqa: E402
from animeface_b200.ops.mbstd import minibatch_stddev                       # noqa: E402
from animeface_b200.ops.resample import Up2xAdjFn, avgpool2, upsample2x_blur  # noqa: E402

DEV, B = 'cuda', 32
SHAPES = {'64': (64, 64, 256), '128': (128, 128, 128), '256': (256, 256, 64), '512': (512, 512, 32), '32': (32, 64, 256),
          '64x32': (64, 32, 256), '32x32': (32, 32, 256)}
DEFAULT = ['mod64x32', 'mod32x32', 'mod128', 'fwd64', 'fwd128', 'dgradpl64', 'dgradpl128', 'wgradpl64', 'wgradpl128', 'wgradpl512',
           'prep64', 'preppool64', 'split64', 'fwdk1', 'up2x_fwd', 'up2x_adj', 'avgpool', 'u4_nhwc', 'u4_nchw', 'u4_down2', 'bias_act', 'mbstd', 'diffaug']


def cl(*shape):
    return torch.randn(*shape, device=DEV).contiguous(memory_format=torch.channels_last)


def twice(fn):
    for _ in range(2):
        fn()


def main():
    which = sys.argv[1:] or DEFAULT
    with torch.no_grad():
        for name in which:
            if name == 'up2x_adj':
                g = cl(B, 64, 256, 256); twice(lambda: Up2xAdjFn.apply(g, True))
            elif name == 'up2x_fwd':
                x = cl(B, 64, 128, 128); twice(lambda: upsample2x_blur(x))
            elif name == 'avgpool':
                x = cl(B, 64, 256, 256); twice(lambda: avgpool2(x))
            elif name in ('u4_nhwc', 'u4_nchw', 'u4_down2'):
                x = cl(B, 64, 256, 256)
                if name == 'u4_nchw':
                    x = x.contiguous()
                f = U.setup_filter([1, 3, 3, 1], device=DEV)
                twice((lambda: U.upfirdn2d(x, f, down=2, padding=1)) if name == 'u4_down2' else (lambda: U.upfirdn2d(x, f, padding=2)))
            elif name == 'bias_act':
                x, b = cl(B, 64, 256, 256), torch.randn(64, device=DEV); twice(lambda: bias_act(x, b, act='lrelu'))
            elif name == 'mbstd':
                x = cl(B, 512, 4, 4); twice(lambda: minibatch_stddev(x, 4))
            elif name == 'diffaug':
                from animeface_b200.diffaugment import DiffAugment
                x = torch.rand(B, 3, 256, 256, device=DEV); twice(lambda: DiffAugment(x, 'color,translation'))
            elif name in ('prep64', 'split64'):
                gy, y = cl(B, 64, 256, 256), cl(B, 64, 256, 256)
                twice((lambda: C._bwd_prep_planes(gy, y, 0.2)) if name == 'prep64' else (lambda: C._split_planes(gy)))
            elif name == 'preppool64':
                gp, y = cl(B, 64, 128, 128), cl(B, 64, 256, 256)
                twice(lambda: C._bwd_prep_planes(gp, y, 0.2, pooled=True, gscale=0.2))
            elif name == 'fwdk1':
                x, w1 = cl(B, 32, 256, 256), torch.randn(64, 32, 1, 1, device=DEV)
                bias = torch.randn(64, device=DEV)
                twice(lambda: C._conv_raw(x, w1, 0.1, False, bias=bias))
            else:
                kind = name.rstrip('0123456789x')
                ci, co, r = SHAPES[name[len(kind):]]
                w = torch.randn(co, ci, 3, 3, device=DEV)
                if kind == 'fwd':
                    x = cl(B, ci, r, r); twice(lambda: C._conv_raw(x, w, 0.1, False))
                elif kind == 'mod':
                    x = cl(B, ci, r, r)
                    s, d = torch.rand(B, ci, device=DEV) + 0.5, torch.rand(B, co, device=DEV) + 0.5
                    bias, nz = torch.randn(co, device=DEV), torch.randn(B, 1, r, r, device=DEV)
                    twice(lambda: C._conv_raw(x, w, 0.1, False, in_scale=s, out_scale=d, bias=bias, noise=nz, slope=0.2))
                elif kind == 'dgrad':
                    x = cl(B, co, r, r); twice(lambda: C._conv_raw(x, w, 0.1, True))
                elif kind == 'wgrad':
                    x, gy = cl(B, ci, r, r), cl(B, co, r, r); twice(lambda: C._wgrad_raw(x, gy, 3, 0.1))
                elif kind == 'dgradpl':
                    gp = C._split_planes(cl(B, co, r, r)); twice(lambda: C._conv_planes(gp, w, 0.1, True))
                elif kind == 'wgradpl':
                    xp, gp = C._split_planes(cl(B, ci, r, r)), C._split_planes(cl(B, co, r, r)); twice(lambda: C._wgrad_planes(xp, gp, 3, 0.1))
                else:
                    raise SystemExit(f'unknown target {name}')
            torch.cuda.synchronize()
            print(name, 'done', flush=True)


if __name__ == '__main__':
    main()
