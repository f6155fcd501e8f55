"""bench.py -- StyleGAN2 256x256 G+D training step, images/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

Workload = BASELINE config 2 ("StyleGAN2 256x256, batch 32, 1xB200, R1 every 16 steps"): defaults of
implementations/StyleGAN2/utils.py:142-160 at image_size 256, fp32 (AMP off -- the parity configuration),
DiffAugment 'color,translation', lazy R1 (d_k = 16), synthetic uniform[-1,1] images, reference init.
A step = Trainer.step = D phase + G phase + EMA (the loop body of utils.py:53-116).
One JSON line on stdout (rank 0).  N > 1: launched by torchrun, one rank per GPU, B = 32 per GPU (weak scaling),
one NCCL all-reduce of the flat gradient buffer per optimizer step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)



TOP_KERNEL = 'halo::conv_halo_kernel<128, 1> grid 148'      # largest single kernel of the step (profiles/r2e_launches.txt)


def load_traffic():
    """(DRAM bytes per launch of the step's largest kernel, the per-kernel table) from the committed `ncu --set full` capture
    (profiles/r4_ncu_traffic.json, written by scripts/ncu_summary.py from the .ncu-rep of scripts/ncu_targets.py: full layer
    shapes, B = 32, dram__bytes_read.sum + dram__bytes_write.sum); (None, None) when the file is absent."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'r4_ncu_traffic.json')))['kernels']
        return t.get(TOP_KERNEL, {}).get('dram_bytes'), {k: v['dram_bytes'] for k, v in t.items()}
    except Exception:
        return None, None


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return dict(hbm=float(p['hbm_gbs']), tf=float(p['bf16_tflops_sustained']), src='measured')
    except Exception:
        return dict(hbm=6650.0, tf=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (recipe's clocks line)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None),
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------
def cpu_oracle_step_rate(batch, steps, warmup, threads=None):
    """img/s of the reference's CPU path (oracle port, plain PyTorch fp32) on this host's cores."""
    import torch
    from oracle import sg2_torch as T
    if threads:
        torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(0)
    sd_g = {k: v.requires_grad_(not k.endswith('.kernel')) for k, v in T.init_generator_sd(gen=gen).items()}
    sd_d = {k: v.requires_grad_(True) for k, v in T.init_discriminator_sd(gen=gen).items()}
    sd_e = {k: v.detach().clone() for k, v in sd_g.items()}
    cfg = T.StepConfig()
    g_lr, g_b, d_lr, d_b = T.adam_hparams(cfg)
    opt_g = torch.optim.Adam([v for v in sd_g.values() if v.requires_grad], lr=g_lr, betas=g_b)
    opt_d = torch.optim.Adam(list(sd_d.values()), lr=d_lr, betas=d_b)
    rng = T.FreshDraws('cpu')
    times = []
    for it in range(warmup + steps):
        real = torch.rand(batch, 3, 256, 256) * 2 - 1
        t0 = time.perf_counter()
        T.train_step(sd_g, sd_d, sd_e, opt_g, opt_d, real, it + 1, rng, cfg)     # it+1: never index 0; R1 when (it+1)%16==0
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * steps / total, total / steps, torch.get_num_threads()


def workload_config(args):
    """The `config` both arms print: BASELINE config 2 (the metric's configuration) or, with --config 4, config 4."""
    if args.config == 4:
        return dict(workload='BASELINE config 4: StyleGAN2 512x512 (channels=32, max 512, style_dim 512) + ADA augment pipeline '
                             '(nnutils.ada.ADA in place of DiffAugment), batch 16 per GPU, R1 every 16 steps, Adam, EMA; fp32 storage',
                    batch_per_gpu=args.batch, image_size=512)
    return dict(workload='BASELINE config 2: StyleGAN2 256x256 (channels=32, max 512, style_dim 512), batch 32 per GPU, '
                         'R1 every 16 steps, DiffAugment color,translation, Adam, EMA; fp32 storage',
                batch_per_gpu=args.batch, image_size=256)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # the reference's step on the host cores costs ~0.4 s per image: a step of the full batch (32) when the run is short
    # enough to end within a few minutes, else a bounded sample of it (stated in `sample`)
    total = args.steps + args.warmup
    batch = args.batch if total <= 6 else (16 if total <= 12 else (8 if total <= 26 else 4))
    batch = min(batch, args.batch)
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers: override it)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    ips, sec, cores = cpu_oracle_step_rate(batch, args.steps, args.warmup, threads=threads)
    sample = (f'oracle port (oracle/sg2_torch.py, plain PyTorch fp32 CPU) of the same step at B={batch} per step '
              f'({"the full batch" if batch == args.batch else "a bounded sample of the batch of " + str(args.batch)}), {args.steps} timed steps')
    cfg = workload_config(args)
    cfg['sample_batch'] = batch
    line = dict(impl='reference', metric='StyleGAN2 256px G+D step images/sec', value=round(ips, 4), unit='images/sec',
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=round(sec * 1e3, 1),
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=cfg,
                cpu_baseline=dict(value=round(ips, 4), unit='images/sec', cores=cores, kind='port', sample=sample),
                e2e=dict(value=round(ips, 4), unit='images/sec', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def upfirdn2d_rates(dev, peaks):
    """Achieved HBM GB/s of the upfirdn2d-family kernels on SURVEY 8(d)'s inputs (algorithmic bytes = one read of x + one
    write of y), CUDA events, L2 flushed between iterations, median of 7."""
    import torch
    from animeface_b200.ops import upfirdn2d as U
    from animeface_b200.ops.resample import Up2xAdjFn, avgpool2, upsample2x_blur
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def rate(fn, nbytes):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        gbs = nbytes / ms / 1e6
        return dict(ms=round(ms, 4), GBps=round(gbs, 1), frac=round(gbs / peaks['hbm'], 3))

    cl = lambda *shape: torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)
    out = {}
    with torch.no_grad():
        x = cl(32, 64, 128, 128)
        out['U1 up2 bilinear+blur [32,64,128,128]->256^2 (fused, SG2 generator)'] = rate(lambda: upsample2x_blur(x), x.numel() * 4 * 5)
        g = cl(32, 64, 256, 256)
        out['U1 adjoint (backward of the fused up2 + blur) [32,64,256,256]->128^2'] = rate(lambda: Up2xAdjFn.apply(g, True), g.numel() * 4 * 1.25)
        out['U3 down2 [1,1] = AvgPool2 [32,64,256,256] (SG2 discriminator)'] = rate(lambda: avgpool2(g), g.numel() * 4 * 1.25)
        f = U.setup_filter([1, 3, 3, 1], device=dev)
        out['U4 filter 4x4 pad 2 -> 257^2, NHWC (SG3-style discriminator)'] = rate(lambda: U.upfirdn2d(g, f, padding=2), (g.numel() + 32 * 64 * 257 * 257) * 4)
        out['U4 down2 4x4 pad 1 -> 128^2, NHWC'] = rate(lambda: U.upfirdn2d(g, f, down=2, padding=1), g.numel() * 4 * 1.25)
        gc = g.contiguous()
        out['U4 filter 4x4 pad 2 -> 257^2, NCHW'] = rate(lambda: U.upfirdn2d(gc, f, padding=2), (g.numel() + 32 * 64 * 257 * 257) * 4)
    out['peak_GBps'] = peaks['hbm']
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from animeface_b200 import _lib
    from animeface_b200.nnutils import init_distributed
    from animeface_b200.ops import conv2d as C
    from animeface_b200.train import GraphedTrainer, TrainConfig, Trainer, build_models, build_optimizers

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device and there is no CPU fallback for the product path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    json_fd = None
    if world > 1:
        # NCCL prints its version banner / debug lines on the C-level stdout; stdout must carry the ONE JSON line of this
        # script.  Point fd 1 at stderr for the rest of the run and keep the original stdout for the JSON line.
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        init_distributed('nccl')
    peaks = load_peaks()
    B = args.batch
    size = 512 if args.config == 4 else 256
    cfg = TrainConfig(batch_size=B, image_size=size, augment='ada' if args.config == 4 else 'diffaugment')
    if args.config == 4:
        args.graphs = False        # the ADA geometry reads its data-dependent padding back to the host (augment.py:281): no capture
    torch.manual_seed(0)                                 # identical replicas (weights), rank-distinct data below
    G, G_ema, D = build_models(cfg, dev)
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    eager = Trainer(cfg, G, G_ema, D, opt_g, opt_d)
    tr = GraphedTrainer(eager) if args.graphs else eager
    torch.manual_seed(1000 + rank)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # resident synthetic batches (uniform [-1,1] like T.Normalize(0.5,0.5) output), a few so steps differ
    pool = [torch.rand(B, 3, size, size, device=dev) * 2 - 1 for _ in range(4)]
    if args.graphs:
        tr.prime(pool[0])                                 # eager + capture of both step kinds (4 untimed steps)
        eager.batches_done = 0
    for i in range(args.warmup):
        tr.step(pool[i % len(pool)])
    sync()

    # ---- device-resident timed region: EXACTLY K steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    C.launch_log = None if args.graphs else []
    launches0 = _lib.launch_count()
    launches_per = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        tr.step(pool[i % len(pool)])
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    conv_log, C.launch_log = C.launch_log, None
    roofline_note = 'timed region'
    if args.config == 4 and eager.ada is not None:
        eager.ada.p.fill_(0.5)     # a mid-training augmentation strength (p starts at 0 = identity transforms)
    if args.graphs:
        # graph replays bypass the host-side launch counter and the per-launch events: measure both in a separate
        # eager pass over the same K steps (same schedule position), AFTER the timed region
        saved_done = eager.batches_done
        eager.batches_done = args.warmup
        C.launch_log = []
        launches0 = _lib.launch_count()
        for i in range(args.steps):
            eager.step(pool[i % len(pool)])
        sync()
        launches = _lib.launch_count() - launches0
        conv_log, C.launch_log = C.launch_log, None
        eager.batches_done = saved_done + args.steps
        roofline_note = 'separate eager pass of the same K steps (the timed region replays CUDA graphs)'

    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = world * B * args.steps / (ms / 1e3)

    # roofline of the dominant kernel family: the convolution launches (fwd/dgrad/wgrad implicit GEMMs)
    conv_ms = sum(a.elapsed_time(b) for _, _, a, b, _ in conv_log)
    conv_flops = sum(f for _, f, _, _, _ in conv_log)
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    by_kind = {}
    by_layer = {}
    for kind, f, a, b, label in conv_log:
        d = by_kind.setdefault(kind, [0.0, 0.0, 0])
        d[0] += f; d[1] += a.elapsed_time(b); d[2] += 1
        d = by_layer.setdefault(f'{kind} {label}', [0.0, 0.0, 0])
        d[0] += f; d[1] += a.elapsed_time(b); d[2] += 1
    if args.layers and rank == 0:
        with open(args.layers, 'w') as fh:
            for k, v in sorted(by_layer.items(), key=lambda kv: -kv[1][1]):
                fh.write(f'{v[1] / args.steps:8.3f} ms/step {100 * v[1] / max(conv_ms, 1e-9):5.1f}% {v[0] / max(v[1], 1e-9) / 1e9:7.1f} TF/s {v[2] / args.steps:5.1f} launches/step  {k}\n')
    roofline = dict(bound='tensor', achieved=round(achieved, 2), peak=peaks['tf'], unit='TFLOP/s',
                    frac=round(achieved / peaks['tf'], 4), traffic=load_traffic()[0], traffic_kernel=TOP_KERNEL + ' (128->128 @128^2, B=32: algorithmic 537 MB)',
                    traffic_by_kernel=load_traffic()[1],
                    kernel='conv2d fwd/dgrad/wgrad (all launches of K steps)', measured_in=roofline_note,
                    peak_source=f"bf16_tflops_sustained of MEASURED_PEAKS.json ({peaks['src']})",
                    share_of_step=round(conv_ms / max(ms, 1e-9), 3),
                    launches=len(conv_log),
                    by_kind={k: dict(tflops=round(v[0] / (v[1] / 1e3) / 1e12, 2), ms=round(v[1], 2), launches=v[2])
                             for k, v in by_kind.items()},
                    mma_per_product=3,
                    note='fp32 operands are split into two 16-bit planes (fp32-class results, DESIGN 3.1): every algorithmic '
                         'product costs 3 tensor-core MMAs, so the tensor pipe does 3x `achieved`; ncu tensor-pipe-active '
                         'and DRAM bytes per launch: profiles/r4_ncu_full.txt')

    # ---- end-to-end: host batch in pinned memory -> H2D each step, losses read back each step
    host = [torch.empty(B, 3, size, size, pin_memory=True).uniform_(-1, 1) for _ in range(2)]
    sync()
    e0.record()
    for i in range(args.steps):
        real = host[i % 2].to(dev, non_blocking=True)
        d_loss, g_loss, _ = tr.step(real)
        _ = (d_loss.item(), g_loss.item())                # D2H read of the step's result
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t)
    e2e = dict(value=round(world * B * args.steps / (e2e_ms / 1e3), 2), unit='images/sec',
               h2d_bytes_per_step=B * 3 * size * size * 4, d2h_bytes_per_step=8)

    def finish():
        # Every rank leaves through here.  The CUDA graphs hold captured NCCL work, and tearing the process group down
        # under them (or with a peer already gone) can block: synchronise, then exit without the teardown.
        sys.stdout.flush()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return
    # second half of BASELINE.json's metric: upfirdn2d achieved HBM GB/s on the SURVEY 8(d) inputs (rank 0, N = 1 only)
    upf = upfirdn2d_rates(dev, peaks) if (world == 1 and args.config == 2) else None
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and args.config == 2:
        ips, sec, cores = cpu_oracle_step_rate(8, 2, 1, threads=(len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else None))
        cpu_baseline = dict(value=round(ips, 4), unit='images/sec', cores=cores, kind='port',
                            sample='oracle/sg2_torch.py (plain-PyTorch CPU restatement of the reference step) on a bounded sample, B=8 of the 32: '
                                   f'1 warm-up + 2 timed steps, {sec:.1f} s/step')
    wcfg = workload_config(args)
    executed_gflop_per_img = conv_flops / (args.steps * B) / 1e9       # convolution flops the step really executes, per image
    wcfg.update(parallelism=f'dp{world}', conv_impl=args.conv_impl, cuda_graphs=bool(args.graphs),
                l2='per-step working set (activations, several GB) >> 126 MB L2; no explicit flush',
                r1_steps_in_timed_region=sum(1 for i in range(args.steps) if (args.warmup + i) % cfg.d_k == 0 and (args.warmup + i) != 0),
                executed_conv_gflop_per_image=round(executed_gflop_per_img, 1),
                model_tflops=round(value * executed_gflop_per_img / 1e3, 2))
    line = dict(metric=f'StyleGAN2 {size}px G+D step images/sec', value=round(value, 2), unit='images/sec', n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=round(ms / args.steps, 3), higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', config=wcfg,
                clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu_baseline,
                upfirdn2d=upf)
    if json_fd is not None:
        os.write(json_fd, (json.dumps(line) + '\n').encode())
    else:
        print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=[2, 4], help='BASELINE config: 2 (256 px, the metric) or 4 (512 px + ADA)')
    ap.add_argument('--batch', type=int, default=None, help='per-GPU batch (BASELINE config 2/3: 32, config 4: 16)')
    ap.add_argument('--conv-impl', default='auto', choices=['auto', 'simt'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--layers', default='', help='write the per-layer convolution time table of the roofline pass to this file')
    ap.add_argument('--no-graphs', dest='graphs', action='store_false', help='run the step eagerly instead of replaying CUDA graphs')
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 16 if args.config == 4 else 32
    if args.impl == 'reference':
        run_reference(args)
        return
    from animeface_b200.ops import conv2d as C
    C.set_default_impl(dict(auto=0, simt=1)[args.conv_impl])
    run_b200(args)


if __name__ == '__main__':
    main()
