"""Kernel micro-benchmarks (CUDA events, L2 flushed between iterations).  Prints one JSON line per kernel.
Usage: python scripts/microbench.py [--quick]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops import conv2d as C                      # noqa: E402
from animeface_b200.ops import upfirdn2d as U                   # noqa: E402
from animeface_b200.ops.bias_act import bias_act                # noqa: E402
from animeface_b200.ops.resample import Up2xAdjFn, avgpool2, upsample2x_blur  # noqa: E402

DEV = 'cuda'
PEAK = 6530.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
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


def report(name, ms, bytes_=None, flops=None):
    r = dict(kernel=name, ms=round(ms, 4))
    if bytes_:
        r['GBps'] = round(bytes_ / ms / 1e6, 1); r['hbm_frac'] = round(bytes_ / ms / 1e6 / PEAK, 3)
    if flops:
        r['TFLOPs'] = round(flops / ms / 1e9, 2)
    print(json.dumps(r), flush=True)


def main():
    quick = '--quick' in sys.argv
    B = 32
    with torch.no_grad():
        # U1: fused bilinear x2 + blur, [32,64,128,128] -> 256^2 (largest SG2-G call)
        x = cl(B, 64, 128, 128)
        report('up2x_blur_fwd U1', timeit(lambda: upsample2x_blur(x)), bytes_=x.numel() * 4 * 5)
        g = cl(B, 64, 256, 256)
        report('up2x_blur_adj U1', timeit(lambda: Up2xAdjFn.apply(g, True)), bytes_=x.numel() * 4 * 5)
        x2 = cl(B, 128, 64, 64)
        report('up2x_blur_fwd U2', timeit(lambda: upsample2x_blur(x2)), bytes_=x2.numel() * 4 * 5)
        # U3: avgpool2 [32,64,256,256]
        report('avgpool2 U3', timeit(lambda: avgpool2(g)), bytes_=g.numel() * 4 * 1.25)
        g2 = cl(B, 64, 256, 256)
        report('avgpool2+residual', timeit(lambda: avgpool2(g, g2, 0.7071)), bytes_=g.numel() * 4 * 2.25)
        del g2
        # U4: generic upfirdn2d, sg3 semantics
        f = U.setup_filter([1, 3, 3, 1], device=DEV)
        report('upfirdn2d U4 filter pad2 nhwc', timeit(lambda: U.upfirdn2d(g, f, padding=2)), bytes_=(g.numel() + B * 64 * 257 * 257) * 4)
        report('upfirdn2d U4 down2 nhwc', timeit(lambda: U.upfirdn2d(g, f, down=2, padding=1)), bytes_=g.numel() * 4 * 1.25)
        report('upfirdn2d up2 nhwc (upsample2d)', timeit(lambda: U.upsample2d(x, f)), bytes_=x.numel() * 4 * 5)
        gc = g.contiguous()
        report('upfirdn2d U4 filter pad2 nchw', timeit(lambda: U.upfirdn2d(gc, f, padding=2)), bytes_=(g.numel() + B * 64 * 257 * 257) * 4)
        report('upfirdn2d U4 down2 nchw', timeit(lambda: U.upfirdn2d(gc, f, down=2, padding=1)), bytes_=g.numel() * 4 * 1.25)
        # bias_act lrelu on [32,64,256,256]
        b = torch.randn(64, device=DEV)
        report('bias_act lrelu nhwc', timeit(lambda: bias_act(g, b, act='lrelu')), bytes_=g.numel() * 8)
        report('bias_act lrelu nchw', timeit(lambda: bias_act(gc, b, act='lrelu')), bytes_=g.numel() * 8)
        # torch copy as the yardstick measured the same way
        dst = torch.empty_like(g)
        report('torch copy_ (yardstick)', timeit(lambda: dst.copy_(g)), bytes_=g.numel() * 8)
        del gc, dst
        # RGB-side 1x1 layers (conv_thin.cu): HBM streams over the wide tensor
        rgb = torch.randn(B, 3, 256, 256, device=DEV)
        wide = cl(B, 32, 256, 256)
        w_in, w_out = torch.randn(32, 3, 1, 1, device=DEV), torch.randn(3, 32, 1, 1, device=DEV)
        bb = torch.randn(32, device=DEV)
        wb = wide.numel() * 4
        report('from_rgb fwd 3->32@256 (+bias+lrelu)', timeit(lambda: C._conv_raw(rgb, w_in, 0.5, False, bias=bb, slope=0.2)), bytes_=wb + rgb.numel() * 4)
        report('from_rgb wgrad 3->32@256', timeit(lambda: C._wgrad_raw(rgb, wide, 1, 0.5)), bytes_=wb + rgb.numel() * 4)
        report('to_rgb fwd 32->3@256 (nchw out)', timeit(lambda: C._conv_raw(wide, w_out, 0.5, False, out_nchw=True)), bytes_=wb + rgb.numel() * 4)
        report('to_rgb dgrad 3->32@256', timeit(lambda: C._conv_raw(rgb, w_out, 0.5, True)), bytes_=wb + rgb.numel() * 4)
        report('to_rgb wgrad 32->3@256', timeit(lambda: C._wgrad_raw(wide, rgb, 1, 0.5)), bytes_=wb + rgb.numel() * 4)
        del rgb, wide
        # convolution layers of the path (fwd / dgrad / wgrad), impl auto
        layers = [(32, 64, 256), (64, 64, 256), (64, 128, 128), (128, 128, 128), (256, 256, 64), (512, 512, 32), (512, 512, 16), (512, 512, 4)]
        if quick:
            layers = layers[1:2] + layers[4:6]
        for ci, co, r in layers:
            xx = cl(B, ci, r, r)
            w = torch.randn(co, ci, 3, 3, device=DEV)
            gy = cl(B, co, r, r)
            fl = 2.0 * B * r * r * ci * co * 9
            byt = (xx.numel() + gy.numel()) * 4
            report(f'conv3x3 fwd {ci}->{co}@{r}', timeit(lambda: C._conv_raw(xx, w, 0.1, False), iters=5, warmup=2), bytes_=byt, flops=fl)
            report(f'conv3x3 dgrad {ci}->{co}@{r}', timeit(lambda: C._conv_raw(gy, w, 0.1, True), iters=5, warmup=2), bytes_=byt, flops=fl)
            report(f'conv3x3 wgrad {ci}->{co}@{r}', timeit(lambda: C._wgrad_raw(xx, gy, 3, 0.1), iters=5, warmup=2), bytes_=byt, flops=fl)
            del xx, gy


if __name__ == '__main__':
    main()
