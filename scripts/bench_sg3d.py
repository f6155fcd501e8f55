"""StyleGAN3-style discriminator at 256 px, B = 16 (BASELINE config 5's discriminator): time of one
forward + backward of the D loss, and of an R1 evaluation, on the libsg2b200 ops."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animeface_b200.nnutils.loss import NonSaturatingLoss, r1_regularizer       # noqa: E402
from animeface_b200.stylegan3 import Discriminator                              # noqa: E402

dev = 'cuda'
torch.manual_seed(0)
D = Discriminator(256).to(dev)
B = 16
real = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
fake = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
loss = NonSaturatingLoss()


def step():
    for p in D.parameters():
        p.grad = None
    loss.d_loss(D(real), D(fake)).backward()


def r1_step():
    for p in D.parameters():
        p.grad = None
    (r1_regularizer()(real, D, None) * 160).backward()


for name, fn in (('D loss fwd+bwd (2 x B=16)', step), ('R1 fwd+double bwd (B=16)', r1_step)):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1) / 3:.1f} ms', flush=True)
