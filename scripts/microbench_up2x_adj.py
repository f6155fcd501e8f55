import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.ops.resample import Up2xAdjFn, upsample2x_blur
g = torch.randn(32, 64, 256, 256, device='cuda').contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def t(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(9):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[4]
with torch.no_grad():
    ms = t(lambda: Up2xAdjFn.apply(g, True))
print(f'up2x_adj [32,64,256,256]: {ms:.4f} ms  {g.numel() * 4 * 1.25 / ms / 1e6:.0f} GB/s')
