"""CPU, world_size 2, gloo: the host-side data-parallel logic (MiniAccelerator + AllReduceOptimizer).
Two ranks with different data must end with identical parameters, equal to one process that sees the mean
gradient -- including a parameter whose gradient is None on one step (zero-filled in the flat buffer)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 1))


def _loss(net, x, it):
    h = net[2](net[1](net[0](x)))
    return (h if it == 1 else net[3](h)).square().mean()      # step 1: the last layer gets no gradient


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from animeface_b200.nnutils import MiniAccelerator
    from animeface_b200.nnutils.accelerate import AllReduceOptimizer, init_distributed
    init_distributed('gloo')
    acc = MiniAccelerator(amp=False, device=torch.device('cpu'))
    assert acc.world_size == world
    net = _net()
    if rank == 1:                                   # perturb: prepare() must broadcast rank 0's weights
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2, betas=(0., 0.99))
    net, opt = acc.prepare(net, opt)
    assert isinstance(opt, AllReduceOptimizer)
    data = torch.Generator().manual_seed(123)
    xs = torch.randn(3, world, 4, 6, generator=data)
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        acc.backward(_loss(net, xs[it, rank], it))
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save(dict(params=gathered, xs=xs), out)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / 'dp.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    a, b = res['params']
    assert torch.equal(a, b), 'ranks diverged'
    # single process, same schedule, gradient = mean over the two ranks' losses
    net = _net()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2, betas=(0., 0.99))
    xs = res['xs']
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        (0.5 * (_loss(net, xs[it, 0], it) + _loss(net, xs[it, 1], it))).backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert torch.allclose(a, ref, rtol=1e-5, atol=1e-6)


def test_single_process_accelerator_api():
    """Reference API surface (nnutils/accelerate.py:134-252): prepare / backward / autocast / update / scaler / device."""
    from animeface_b200.nnutils import MiniAccelerator
    acc = MiniAccelerator(amp=False, device=torch.device('cpu'))
    net = _net()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    net2, opt2, other = acc.prepare(net, opt, 'passthrough')
    assert net2 is net and opt2 is opt and other == 'passthrough'
    assert acc.scaler is None and acc.device == torch.device('cpu')
    with acc.autocast():
        loss = net(torch.zeros(2, 6)).sum()
    acc.backward(loss)
    opt2.step()
    acc.update()
    assert acc.prepare() is None
