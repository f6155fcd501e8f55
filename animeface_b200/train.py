"""The StyleGAN2 G+D training step of the reference on the B200 kernels.

Reference: the loop body of implementations/StyleGAN2/utils.py:53-116 (``train``) and the optimizer set-up
:208-221.  Same losses, same lazy-R1 schedule (the GAN loss is DROPPED on R1 steps, :71-76), same Adam
hyper-parameters, same random-draw order.  What is removed from the timed step are the reference's host
round trips: ``save_image`` every step (:124), ``.item()`` / ``isnan().any()`` syncs (:127-130); losses stay on
the device.  Two equivalences are used (results identical, work skipped):
  * D phase: G runs under no_grad (the reference builds the graph and detaches, :67-69);
  * G phase: D's parameters do not require grad (the reference computes their gradients and then discards them
    at the next ``zero_grad``, :55-56).
Path-length regularisation (:18-33, 96-103; off by default in the reference, pl_lambda = 0 at :159) runs on steps
``batches_done % g_k == 0``: the generator forward of that step is built from the any-order modulated convolution
(ops.conv2d.any_order_modconv) so that d(sum(img * noise))/d style can itself be differentiated; like the reference
the GAN loss of the generator is dropped on those steps, and the running ``pl_mean`` is the EMA(0.99) of the PENALTY
value (utils.py:100-103) -- kept as a device scalar here instead of the reference's ``.cpu().numpy()`` round trip.
"""
from __future__ import annotations

import contextlib
import functools
import math
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import rng
from .ops.conv2d import any_order_modconv
from .diffaugment import DiffAugment
from .model import Discriminator, Generator, independent_batches, init_weight_N01, supplied_noise  # noqa: F401
from .nnutils import FlatAdam, init_distributed, update_ema
from .nnutils.loss import NonSaturatingLoss, calc_grad, r1_regularizer


def pl_penalty(styles, images, pl_mean, scaler=None):
    """Path-length regulariser (implementations/StyleGAN2/utils.py:18-29): with noise ~ N(0,1)/sqrt(H*W),
    g = d sum(images * noise) / d styles (create_graph), penalty = mean_b((||g_b||_2 - pl_mean)^2).
    ``images`` must come from a generator forward run under ``any_order_modconv``."""
    num_pixels = images.size()[2:].numel()
    noise = rng.randn(*images.size(), device=images.device) / math.sqrt(num_pixels)
    outputs = (images * noise).sum()
    gradients = calc_grad(outputs, styles, scaler)
    gradients = gradients.pow(2).sum(dim=1).sqrt()
    return (gradients - pl_mean).pow(2).mean()


def update_pl_mean(old, new, decay=0.99):
    """exponential moving average (implementations/StyleGAN2/utils.py:31-33)"""
    return decay * old + (1 - decay) * new


@dataclass
class TrainConfig:
    """Defaults of implementations/StyleGAN2/utils.py:142-160 + utils/argument.py:10-31 at 256 px."""
    image_size: int = 256
    image_channels: int = 3
    style_dim: int = 512
    channels: int = 32
    max_channels: int = 512
    block_num_conv: int = 2
    map_num_layers: int = 8
    map_lr: float = 0.01
    mbsd_groups: int = 4
    batch_size: int = 32
    lr: float = 1e-3
    beta1: float = 0.
    beta2: float = 0.99
    g_k: int = 8
    d_k: int = 16
    r1_lambda: float = 10.
    pl_lambda: float = 0.
    policy: str = 'color,translation'
    ema_decay: float = 0.999
    # 'diffaugment' = the StyleGAN2 loop's DiffAugment(policy) (utils.py:63-68); 'ada' = BASELINE config 4: nnutils.ada.ADA(batch_size)
    # in its place, its strength p following sign(D(real)) (nnutils/ada.py:25-36; implementations/ADA/utils.py:70)
    augment: str = 'diffaugment'


def build_models(cfg: TrainConfig, device):
    """Models + init exactly as utils.py:186-205.  Under torchrun (one process per GPU) the process group is initialised
    here, rank 0's freshly initialised weights are broadcast so the replicas start identical whatever each rank's seed is,
    and every rank then moves its generators to a rank-distinct stream (latents, noise and augmentation draws must differ
    between ranks -- identical draws would make the global batch `world` copies of one batch)."""
    mk_g = lambda: Generator(cfg.image_size, cfg.image_channels, cfg.style_dim, cfg.channels, cfg.max_channels,
                             cfg.block_num_conv, cfg.map_num_layers, True, cfg.map_lr)
    G, G_ema = mk_g(), mk_g()
    D = Discriminator(cfg.image_size, cfg.image_channels, cfg.channels, cfg.max_channels, cfg.block_num_conv, cfg.mbsd_groups)
    G.init_weight(map_init_func=functools.partial(init_weight_N01, lr=cfg.map_lr), syn_init_func=init_weight_N01)
    G_ema.eval()
    G_ema.load_state_dict(G.state_dict())       # == update_ema(G, G_ema, decay=0) on finite memory (utils.py:199-200)
    D.apply(init_weight_N01)
    G, G_ema, D = G.to(device), G_ema.to(device), D.to(device)
    init_distributed()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():
            for m in (G, G_ema, D):
                for t in list(m.parameters()) + list(m.buffers()):
                    dist.broadcast(t.data, src=0)
        seed = torch.initial_seed() + 7919 * (dist.get_rank() + 1)
        torch.manual_seed(seed)                     # seeds the CPU and every CUDA generator of this process
    return G, G_ema, D


def build_optimizers(cfg: TrainConfig, G, G_ema, D):
    """Lazy-regularisation-scaled Adam (utils.py:208-221) as two FlatAdam instances."""
    betas = (cfg.beta1, cfg.beta2)
    if cfg.pl_lambda > 0:
        r = cfg.g_k / (cfg.g_k + 1)
        g_lr, g_betas = cfg.lr * r, (betas[0] ** r, betas[1] ** r)
    else:
        g_lr, g_betas = cfg.lr, betas
    if cfg.r1_lambda > 0:
        r = cfg.d_k / (cfg.d_k + 1)
        d_lr, d_betas = cfg.lr * r, (betas[0] ** r, betas[1] ** r)
    else:
        d_lr, d_betas = cfg.lr, betas
    opt_g = FlatAdam(G.parameters(), lr=g_lr, betas=g_betas, model=G, ema_model=G_ema)
    opt_d = FlatAdam(D.parameters(), lr=d_lr, betas=d_betas, model=D)
    return opt_g, opt_d


class Trainer:
    """One object = the state of the reference's ``train()`` loop; ``step(real)`` = one iteration."""

    def __init__(self, cfg: TrainConfig, G, G_ema, D, opt_g, opt_d):
        self.cfg, self.G, self.G_ema, self.D, self.opt_g, self.opt_d = cfg, G, G_ema, D, opt_g, opt_d
        self.loss = NonSaturatingLoss()
        self.r1 = r1_regularizer()
        if cfg.augment == 'ada':
            from .ada import ADA
            self.ada = ADA(batch_size=cfg.batch_size).to(next(G.parameters()).device)
            self.augment = self.ada
        else:
            assert cfg.augment == 'diffaugment', cfg.augment
            self.ada = None
            self.augment = functools.partial(DiffAugment, policy=cfg.policy)
        self.batches_done = 0
        self._d_params = list(D.parameters())
        # data parallel: the discriminator's gradient exchange + Adam step run on a side stream while the main stream already
        # computes the generator forward of the G phase (which does not read D); the main stream joins before D is used again.
        # Inside a captured step the fork / join become graph dependencies.
        self._side = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and next(D.parameters()).is_cuda:
            self._side = torch.cuda.Stream(device=next(D.parameters()).device)
        # running mean of the path-length penalty (utils.py:44, 100-103), on the device so the step never syncs
        self.pl_mean = torch.zeros((), dtype=torch.float32, device=next(G.parameters()).device)

    def is_r1_step(self, it=None):
        it = self.batches_done if it is None else it
        return it % self.cfg.d_k == 0 and self.cfg.r1_lambda > 0 and it != 0

    def is_pl_step(self, it=None):
        it = self.batches_done if it is None else it
        return it % self.cfg.g_k == 0 and self.cfg.pl_lambda > 0 and it != 0

    def step(self, real: torch.Tensor, force_r1=None, force_pl=None):
        """real: [B,3,H,W] on the device.  Returns (D_loss, G_loss, fake) device tensors; no host sync.
        force_r1 / force_pl override the lazy-regularisation schedule for this call (used to warm up / capture graphs)."""
        cfg, G, D = self.cfg, self.G, self.D
        B, dev = real.size(0), real.device
        self.opt_g.zero_grad()
        self.opt_d.zero_grad()
        # ---- discriminator phase (utils.py:60-86)
        z = rng.randn(B, cfg.style_dim, device=dev)
        r1_step = self.is_r1_step() if force_r1 is None else bool(force_r1)
        pl_step = self.is_pl_step() if force_pl is None else bool(force_pl)
        real_aug = self.augment(real)
        with torch.no_grad():
            fake, _ = G(z)
            fake_aug = self.augment(fake)
        if r1_step:
            # lazy R1 (utils.py:71-76): the GAN loss is dropped, so D(real_aug) / D(fake_aug) of :65,:69 are dead
            # values -- only their random draws (consumed above) matter for the stream.
            D_loss = self.r1(real, D, None) * cfg.r1_lambda * cfg.d_k
        else:
            # D(real_aug) and D(fake_aug) (utils.py:65,69) as one call over the concatenated batch: same per-sample
            # results (minibatch-stddev is evaluated per half), half the launches and weight packs of the D phase
            with independent_batches(2):
                prob = D(torch.cat([real_aug, fake_aug], dim=0))
            D_loss = self.loss.d_loss(prob[:B], prob[B:])
            if self.ada is not None:
                self.ada.update_p(prob[:B].detach())          # overfitting heuristic on D(real_aug); device-side, no sync
        D_loss.backward()
        d_done = None
        if self._side is not None:
            fork = torch.cuda.current_stream().record_event()
            with torch.cuda.stream(self._side):
                self._side.wait_event(fork)
                self.opt_d.step()
                d_done = self._side.record_event()
        else:
            self.opt_d.step()
        join_d = (lambda: torch.cuda.current_stream().wait_event(d_done)) if d_done is not None else (lambda: None)
        # ---- generator phase (utils.py:88-113)
        z = rng.randn(B, cfg.style_dim, device=dev)
        for p in self._d_params:
            p.requires_grad_(False)
        try:
            if pl_step:
                # lazy path-length step (utils.py:96-103): G_loss is the penalty alone; D(fake_aug) of :93-94 is a dead
                # value, its augmentation draws are still consumed in the reference's order.
                with any_order_modconv():
                    fake, style = G(z)
                join_d()
                with torch.no_grad():
                    self.augment(fake)
                pl = pl_penalty(style, fake, self.pl_mean)
                G_loss = pl * cfg.pl_lambda * cfg.g_k
                G_loss.backward()
                self.pl_mean.copy_(update_pl_mean(self.pl_mean, pl.detach()))
            else:
                fake, _ = G(z)
                join_d()                                    # D's updated weights are needed from here on
                fake_prob = D(self.augment(fake))
                G_loss = self.loss.g_loss(fake_prob)
                G_loss.backward()
        finally:
            for p in self._d_params:
                p.requires_grad_(True)
        self.opt_g.step()
        update_ema(G, self.G_ema, cfg.ema_decay)
        self.batches_done += 1
        return D_loss.detach(), G_loss.detach(), fake.detach()


class GraphedTrainer:
    """``Trainer.step`` captured into CUDA graphs -- one per step kind (normal, lazy-R1, and the path-length variants
    when pl_lambda > 0) -- and replayed.

    A step launches ~1300 kernels (convolutions, resampling, elementwise, optimizer); replaying a graph removes the
    Python / launch gaps between them (guide rule: "capture launch-bound inner loops in CUDA graphs").  Everything in
    the step is capture-safe: no host syncs, random draws from the device generator, tensor maps and kernel arguments
    are functions of addresses that the graph's private memory pool keeps fixed, the Adam step counters and the
    path-length running mean live on the device; under torchrun the NCCL all-reduce of the flat gradient buffer is
    captured with the rest.  Policy: the first step of each kind runs eagerly (it also warms kernels up), the second
    one is captured and replayed, later ones are replays.  ``prime()`` does that up front for every kind of the
    schedule.  A batch of another size (ragged last batch) runs the eager step; a change of lr / betas / eps drops the
    captured graphs (they hold those values as kernel scalars)."""

    def __init__(self, trainer: Trainer, eager_first: bool = True):
        self.t = trainer
        self.eager_first = eager_first      # False: capture a kind at its first occurrence (caller has warmed the kernels up)
        self.static_real = None
        self.graphs = {}          # kind (is_r1, is_pl) -> (CUDAGraph, outputs)
        self.seen = set()
        self._hparams = None

    def kinds(self):
        """Step kinds of the lazy-regularisation schedule, most frequent first."""
        t, out = self.t, []
        for it in range(1, 2 * t.cfg.d_k * t.cfg.g_k + 1):
            k = (t.is_r1_step(it), t.is_pl_step(it))
            if k not in out:
                out.append(k)
        return out

    def prime(self, real):
        """Eager + capture for every kind now (2 extra optimizer steps per kind); the schedule position is kept."""
        done = self.t.batches_done
        for kind in self.kinds():
            self.step(real, *kind)
            self.step(real, *kind)
        self.t.batches_done = done

    def _current_hparams(self):
        # lr / betas / eps are kernel scalars baked into a captured graph: a change must re-capture
        return tuple((g['lr'], tuple(g['betas']), g['eps']) for o in (self.t.opt_g, self.t.opt_d) for g in o.param_groups)

    def step(self, real, force_r1=None, force_pl=None):
        t = self.t
        kind = (t.is_r1_step() if force_r1 is None else bool(force_r1),
                t.is_pl_step() if force_pl is None else bool(force_pl))
        hp = self._current_hparams()
        if self._hparams != hp:
            self.graphs.clear()
            self._hparams = hp
        if self.static_real is None:
            self.static_real = real.clone()
        if real.shape != self.static_real.shape:
            return t.step(real, *kind)          # ragged last batch of a loader: the eager step takes any batch size
        if kind in self.graphs:
            self.static_real.copy_(real)
            g, out = self.graphs[kind]
            g.replay()
            t.batches_done += 1
            return out
        if self.eager_first and kind not in self.seen:
            self.seen.add(kind)
            return t.step(real, *kind)
        self.static_real.copy_(real)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        done = t.batches_done
        with torch.cuda.graph(g):
            out = t.step(self.static_real, *kind)
        t.batches_done = done               # capture does not execute: the replay below is the real step
        self.graphs[kind] = (g, out)
        g.replay()
        t.batches_done += 1
        return out


def train(max_iter, dataset, cfg: TrainConfig, device, log_every=100, log=print):
    """Minimal counterpart of the reference ``train()`` (utils.py:35-138): iterate a loader of real batches.
    Under torchrun every rank runs this function: replicas are synchronised in ``build_models``, gradients in
    ``FlatAdam.step``; a ``DataLoader`` is sharded by rank (``MiniAccelerator.prepare``), any other iterable is taken
    as already rank-local."""
    G, G_ema, D = build_models(cfg, device)
    opt_g, opt_d = build_optimizers(cfg, G, G_ema, D)
    trainer = Trainer(cfg, G, G_ema, D, opt_g, opt_d)
    if isinstance(dataset, torch.utils.data.DataLoader):
        from .nnutils import MiniAccelerator
        dataset = MiniAccelerator(amp=False, device=device).prepare(dataset)
    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    log = log if rank0 else (lambda *a, **k: None)
    while trainer.batches_done < max_iter:
        for real in dataset:
            d_loss, g_loss, _ = trainer.step(real.to(device, non_blocking=True))
            if log_every and trainer.batches_done % log_every == 0:
                log(f'{trainer.batches_done}/{max_iter} D={d_loss.item():.4f} G={g_loss.item():.4f}')
            if trainer.batches_done >= max_iter:
                break
    return trainer
