"""Per-kernel device time of the training step (BASELINE config 2, B = 32), from CUPTI through torch.profiler.
    python scripts/step_profile.py [steps=16] > profiles/<round>_step_kernels.txt
Runs the EAGER Trainer (the CUDA graphs of the bench replay exactly these launches), 3 warm-up steps, then `steps` profiled
steps (16 = one lazy-R1 period).  Kernel times are concurrent-free (one stream), so their sum is the device-busy time of the
step; the table says where the non-convolution share goes.  Unlike an ncu launch list it runs at full clocks and warm caches."""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.train import TrainConfig, Trainer, build_models, build_optimizers      # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device('cuda', 0)
    cfg = TrainConfig(batch_size=32, image_size=256)
    torch.manual_seed(0)
    G, G_ema, D = build_models(cfg, dev)
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    tr = Trainer(cfg, G, G_ema, D, opt_g, opt_d)
    pool = [torch.rand(32, 3, 256, 256, device=dev) * 2 - 1 for _ in range(4)]
    for i in range(3):
        tr.step(pool[i % 4])
    tr.batches_done = 3
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(steps):
            tr.step(pool[i % 4])
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            a = agg[ev.name]
            a[0] += ev.device_time_total if hasattr(ev, 'device_time_total') else ev.cuda_time_total
            a[1] += 1
    total = sum(v[0] for v in agg.values())
    print(f'# {steps} eager steps, {total / 1e3 / steps:.3f} ms of kernel time per step, {sum(v[1] for v in agg.values()) / steps:.0f} launches per step')
    print('#   ms/step  share  launches/step  kernel')
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if us / total < 0.0005:
            continue
        print(f'{us / 1e3 / steps:10.3f} {100 * us / total:5.1f}% {n / steps:8.1f}  {name[:150]}')


if __name__ == '__main__':
    main()
