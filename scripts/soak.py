"""Soak run: N training steps replayed from the CUDA graphs (B = 32, 256 px, synthetic data), losses read back every 50
steps -- checks that the step stays finite and keeps moving over a few hundred optimizer updates, R1 steps included."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.train import GraphedTrainer, TrainConfig, Trainer, build_models, build_optimizers   # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device('cuda')
cfg = TrainConfig(batch_size=32)
torch.manual_seed(0)
G, G_ema, D = build_models(cfg, dev)
opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
gt = GraphedTrainer(Trainer(cfg, G, G_ema, D, opt_g, opt_d))
pool = [torch.rand(32, 3, 256, 256, device=dev) * 2 - 1 for _ in range(8)]
gt.prime(pool[0])
gt.t.batches_done = 0
p0 = opt_g.flat_params.clone()
for it in range(steps):
    d_loss, g_loss, fake = gt.step(pool[it % len(pool)])
    if (it + 1) % 50 == 0 or it == 0:
        ok = bool(torch.isfinite(d_loss) and torch.isfinite(g_loss) and torch.isfinite(fake).all())
        print(f'step {it + 1:4d}: D {float(d_loss):9.4f}  G {float(g_loss):9.4f}  |fake| max {float(fake.abs().max()):.3f}  '
              f'finite {ok}  |dG| {float((opt_g.flat_params - p0).abs().max()):.4f}', flush=True)
        assert ok, 'non-finite value in the step'
print('soak ok; peak memory %.1f GiB' % (torch.cuda.max_memory_allocated() / 2**30))
