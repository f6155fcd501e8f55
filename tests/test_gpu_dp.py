"""GPU, world_size 2 (skipped on a single-GPU box): FlatAdam's fused step with its NCCL all-reduce gives every rank the
same parameters as one process training on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from animeface_b200.nnutils import FlatAdam, MiniAccelerator
    acc = MiniAccelerator(amp=False)
    net = acc.prepare(_net())
    opt = acc.prepare(FlatAdam(net.parameters(), lr=1e-2, betas=(0., 0.99), model=net))
    xs = torch.randn(3, world, 4, 6, generator=torch.Generator().manual_seed(5)).to(acc.device)
    for it in range(3):
        opt.zero_grad()
        net(xs[it, rank]).square().mean().backward()
        opt.step()
    flat = opt.flat_params.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save(dict(params=[g.cpu() for g in gathered], xs=xs.cpu()), out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_flat_adam_two_ranks(tmp_path):
    out = str(tmp_path / 'dp.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    a, b = res['params']
    assert torch.equal(a, b)
    net = _net().cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2, betas=(0., 0.99))
    xs = res['xs'].cuda()
    for it in range(3):
        opt.zero_grad()
        (0.5 * (net(xs[it, 0]).square().mean() + net(xs[it, 1]).square().mean())).backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    got = torch.cat([a[o:o + n] for o, n in ((0, 30), (32, 5), (40, 5), (48, 1))])   # 16-byte aligned segments
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
