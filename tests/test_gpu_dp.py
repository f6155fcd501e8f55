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


# ---------------------------------------------------------------------------------------------------------------------
# The real (small) G/D under data parallelism: SURVEY 8(e) -- one NCCL all-reduce of the flat gradient buffer per
# optimizer step, replicas identical afterwards.

def _sg2_cfg():
    from animeface_b200.train import TrainConfig
    return TrainConfig(image_size=32, style_dim=64, channels=8, max_channels=64, map_num_layers=2, batch_size=8, d_k=2)


def _local_grad_d(tr, real, seed):
    """Flat D gradient of ONE rank's batch at the current weights (no optimizer step): the D phase of Trainer.step with the
    same draws in the same order (latent, augmentation of real, generator noise, augmentation of fake)."""
    from animeface_b200.diffaugment import DiffAugment
    from animeface_b200.model import independent_batches
    torch.manual_seed(seed)
    cfg, G, D = tr.cfg, tr.G, tr.D
    B = real.size(0)
    z = torch.randn(B, cfg.style_dim, device=real.device)
    real_aug = DiffAugment(real, cfg.policy)
    with torch.no_grad():
        fake, _ = G(z)
        fake_aug = DiffAugment(fake, cfg.policy)
    with independent_batches(2):
        prob = D(torch.cat([real_aug, fake_aug], dim=0))
    d_loss = tr.loss.d_loss(prob[:B], prob[B:])
    gd = torch.autograd.grad(d_loss, list(D.parameters()), allow_unused=True)
    return torch.cat([(g if g is not None else torch.zeros_like(p)).reshape(-1) for g, p in zip(gd, D.parameters())])


def _sg2_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    from animeface_b200.train import GraphedTrainer, Trainer, build_models, build_optimizers
    cfg = _sg2_cfg()
    dev = torch.device('cuda', rank)
    torch.manual_seed(100 + rank)                 # DIFFERENT seeds: build_models must broadcast rank 0's weights
    G, G_ema, D = build_models(cfg, dev)          # initialises the process group
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    tr = Trainer(cfg, G, G_ema, D, opt_g, opt_d)
    start = [opt_g.flat_params.clone(), opt_d.flat_params.clone()]
    reals = torch.rand(4, world, 8, 3, 32, 32, generator=torch.Generator().manual_seed(3)).to(dev) * 2 - 1
    # (1) the all-reduced gradient buffer == the sum of the two single-rank gradients (Adam scales it by 1/world)
    gd_local = _local_grad_d(tr, reals[0, rank], 500 + rank)
    torch.manual_seed(500 + rank)                 # the same draws again for the real step
    tr.step(reals[0, rank])
    gathered_d = [torch.zeros_like(gd_local) for _ in range(world)]
    dist.all_gather(gathered_d, gd_local)
    red_d = torch.cat([opt_d.flat_grads[o:o + n] for o, n in zip(opt_d._offs, opt_d._sizes)])
    # (2) three more steps through CUDA graphs with the NCCL all-reduce captured inside (step index 2 is an R1 step)
    gt = GraphedTrainer(tr)
    gt.prime(reals[1, rank])
    for it in (1, 2, 3):
        gt.step(reals[it, rank])
    torch.cuda.synchronize()
    flats = [opt_g.flat_params, opt_d.flat_params, G_ema._sg2_flat]
    peers = []
    for f in flats + start:
        g = [torch.zeros_like(f) for _ in range(world)]
        dist.all_gather(g, f.contiguous())
        peers.append([t.cpu() for t in g])
    if rank == 0:
        torch.save(dict(red_d=red_d.cpu(), loc_d=[t.cpu() for t in gathered_d], peers=peers, r1_seen=sorted(gt.graphs)), out)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)        # graphs hold captured NCCL work: leave without tearing the process group down (as bench.py does)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_stylegan2_step_two_ranks(tmp_path):
    """2 ranks, real small G/D: replicas start identical although the ranks seed differently, the reduced gradient is the
    sum of the per-rank gradients, and after eager + graphed steps (incl. an R1 step) the weights are bit-identical."""
    out = str(tmp_path / 'sg2dp.pt')
    ctx = mp.spawn(_sg2_worker, args=(2, _free_port(), out), nprocs=2, join=False)
    ctx.join()
    res = torch.load(out)
    red, loc = res['red_d'], res['loc_d']
    want = loc[0] + loc[1]
    scale = float(want.abs().max())
    assert scale > 0 and float((loc[0] - loc[1]).abs().max()) > 1e-3 * scale        # the ranks really saw different batches
    assert float((red - want).abs().max()) <= 2e-5 * scale, float((red - want).abs().max()) / scale
    for i, pair in enumerate(res['peers']):
        assert torch.equal(pair[0], pair[1]), f'flat buffer {i} differs between the ranks'
    for after, before in zip(res['peers'][:2], res['peers'][3:5]):
        assert not torch.equal(after[0], before[0])             # ... and the steps really trained
    assert res['r1_seen'] == [(False, False), (True, False)]
