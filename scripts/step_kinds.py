"""Time of each step kind (normal / lazy-R1 / path-length) replayed from its CUDA graph, B = 32, 256 px."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.train import GraphedTrainer, TrainConfig, Trainer, build_models, build_optimizers   # noqa: E402

dev = torch.device('cuda')
pl = float(os.environ.get('PL_LAMBDA', '0'))
cfg = TrainConfig(batch_size=32, pl_lambda=pl)
torch.manual_seed(0)
G, G_ema, D = build_models(cfg, dev)
opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
gt = GraphedTrainer(Trainer(cfg, G, G_ema, D, opt_g, opt_d))
real = torch.rand(32, 3, 256, 256, device=dev) * 2 - 1
gt.prime(real)
for kind in gt.kinds():
    for _ in range(2):
        gt.step(real, *kind)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        gt.step(real, *kind)
    e1.record()
    torch.cuda.synchronize()
    print(f'kind (r1={kind[0]}, pl={kind[1]}): {e0.elapsed_time(e1) / 5:.2f} ms per step', flush=True)
print(f'peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
