"""FlatAdam: torch.optim.Adam semantics over one flat fp32 buffer, one kernel, one all-reduce.

Replaces ``optim.Adam(G.parameters(), ...)`` / ``optim.Adam(D.parameters(), ...)`` of
implementations/StyleGAN2/utils.py:208-221 (and, with ``ema_model=``, the 81 per-tensor lerps of
nnutils/training.py:23-40 become one kernel in ``update_ema``).

  * parameters are re-pointed to views of ONE flat buffer; m / v are flat too;
  * ``step()`` packs the gradients with one multi-tensor copy, all-reduces the flat buffer ONCE when
    torch.distributed is initialised (the only collective of the data-parallel step; there is no DDP in the
    reference, nnutils/accelerate.py:8-11), and launches ``sg2_adam_multi`` once over the whole buffer (per-tensor
    step counts live on the device, so the call replays unchanged inside a CUDA graph);
  * tensors whose grad is None are skipped exactly like torch.optim.Adam skips them (their step count does
    not advance) -- the 12 ``InjectNoise.scale`` and, on R1 steps, the last bias of D.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import _lib


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, model=None, ema_model=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        ps = self.param_groups[0]['params']
        if len(self.param_groups) != 1:
            raise ValueError('FlatAdam takes a single parameter group')
        dev = ps[0].device
        if not all(p.dtype == torch.float32 and p.device == dev for p in ps):
            raise ValueError('FlatAdam: float32 parameters on one device only')
        self._ps = ps
        self._sizes = [p.numel() for p in ps]
        # 16-byte aligned segments so every tensor view stays vector-load friendly
        self._offs, off = [], 0
        for n in self._sizes:
            self._offs.append(off)
            off += (n + 3) // 4 * 4
        self._total = off
        self._P = torch.zeros(off, dtype=torch.float32, device=dev)
        self._G = torch.zeros_like(self._P)
        self._M = torch.zeros_like(self._P)
        self._V = torch.zeros_like(self._P)
        self._steps = torch.zeros(len(ps), dtype=torch.int64, device=dev)
        self._present = torch.zeros(len(ps), dtype=torch.int32, device=dev)
        self._coef = torch.zeros(2 * len(ps), dtype=torch.float32, device=dev)
        for p, o, n in zip(ps, self._offs, self._sizes):
            self._P[o:o + n].copy_(p.data.reshape(-1))
            p.data = self._P[o:o + n].view(p.shape)
        self._seg_off = torch.tensor(self._offs + [self._total], dtype=torch.int64, device=dev)
        self._gviews = [self._G[o:o + n].view(p.shape) for p, o, n in zip(ps, self._offs, self._sizes)]
        if model is not None:
            model._sg2_flat = self._P
        if ema_model is not None:
            self._flatten_like(ema_model, model)

    def _flatten_like(self, ema_model, model):
        """Give ema_model a flat buffer with the same segment layout as the optimised model."""
        if model is None:
            raise ValueError('ema_model needs model=')
        src = dict(model.named_parameters())
        index = {id(p): i for i, p in enumerate(self._ps)}
        E = torch.zeros_like(self._P)
        for key, pe in ema_model.named_parameters():
            i = index[id(src[key])]
            o, n = self._offs[i], self._sizes[i]
            E[o:o + n].copy_(pe.data.reshape(-1))
            pe.data = E[o:o + n].view(pe.shape)
        ema_model._sg2_flat = E

    @property
    def flat_params(self):
        return self._P

    @property
    def flat_grads(self):
        return self._G

    # ---- checkpointing: torch.optim.Adam's layout (state[i] = {step, exp_avg, exp_avg_sq}), so FlatAdam and
    # torch.optim.Adam checkpoints are interchangeable and a resume restores the moments and the bias correction.
    def state_dict(self):
        steps = self._steps.tolist()
        state = {}
        for i, (o, n, p) in enumerate(zip(self._offs, self._sizes, self._ps)):
            if steps[i] == 0:
                continue                                   # torch.optim.Adam has no entry for a never-stepped tensor
            state[i] = dict(step=torch.tensor(float(steps[i])),
                            exp_avg=self._M[o:o + n].view(p.shape).clone(),
                            exp_avg_sq=self._V[o:o + n].view(p.shape).clone())
        group = {k: v for k, v in self.param_groups[0].items() if k != 'params'}
        group['params'] = list(range(len(self._ps)))
        return dict(state=state, param_groups=[group])

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        groups = state_dict['param_groups']
        if len(groups) != 1 or len(groups[0]['params']) != len(self._ps):
            raise ValueError('FlatAdam.load_state_dict: the checkpoint has a different parameter list')
        for k, v in groups[0].items():
            if k != 'params' and k in self.param_groups[0]:
                self.param_groups[0][k] = v
        self._M.zero_(); self._V.zero_()
        steps = [0] * len(self._ps)
        for key, st in state_dict['state'].items():
            i = int(key)
            o, n, p = self._offs[i], self._sizes[i], self._ps[i]
            if tuple(st['exp_avg'].shape) != tuple(p.shape):
                raise ValueError(f'FlatAdam.load_state_dict: state {i} has shape {tuple(st["exp_avg"].shape)}, parameter {tuple(p.shape)}')
            self._M[o:o + n].copy_(st['exp_avg'].reshape(-1))
            self._V[o:o + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps[i] = int(float(st['step']))
        self._steps.copy_(torch.tensor(steps, dtype=torch.int64))

    def _check_views(self):
        """The parameters must still be the views of the flat buffer made in __init__: ``model.to(...)``, ``.half()`` or a
        ``p.data = ...`` afterwards would leave the optimizer updating a buffer nobody reads."""
        base = self._P.data_ptr()
        for p, o in zip(self._ps, self._offs):
            if p.data_ptr() != base + 4 * o:
                raise RuntimeError('FlatAdam: a parameter no longer aliases the flat buffer (model moved / re-typed / .data '
                                   're-assigned after the optimizer was built); build the optimizer last')

    def zero_grad(self, set_to_none: bool = True):
        for p in self._ps:
            p.grad = None
        self._reduced = False

    @torch.no_grad()
    def reduce_gradients(self):
        """Pack the gradients into the flat buffer, all-reduce it (mean over ranks) and re-point every ``p.grad`` at its view
        of the reduced buffer.  ``step()`` then skips its own exchange.  This is the order a GradScaler needs
        (implementations/StyleGAN2/utils.py:85,112 run ``optimizer.step()`` through the scaler): the ranks must agree on the
        scaler's inf check, so it has to look at the REDUCED gradients (``MiniAccelerator`` wires this up when amp=True)."""
        present = [i for i, p in enumerate(self._ps) if p.grad is not None]
        if not present or getattr(self, '_reduced', False):
            return
        self._check_views()
        torch._foreach_copy_([self._gviews[i] for i in present], [self._ps[i].grad for i in present])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self._G)
            self._G.mul_(1.0 / dist.get_world_size())
        for i in present:
            self._ps[i].grad = self._gviews[i]
        self._reduced = True

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        group = self.param_groups[0]
        lr, (b1, b2), eps = group['lr'], group['betas'], group['eps']
        present = [i for i, p in enumerate(self._ps) if p.grad is not None]
        if not present:
            return
        self._check_views()
        scale = 1.0
        if getattr(self, '_reduced', False):
            self._reduced = False                         # reduce_gradients() already packed / exchanged / averaged
        else:
            torch._foreach_copy_([self._gviews[i] for i in present], [self._ps[i].grad for i in present])
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(self._G)                  # the one collective of the step
                scale = 1.0 / dist.get_world_size()
        lib = _lib.load()
        st = _lib.stream_ptr(self._P)
        # presence flags from plain integer ranges (no host->device copy: the call replays inside a CUDA graph)
        self._present.zero_()
        run = [present[0], present[0]]
        for i in present[1:] + [None]:
            if i is not None and i == run[1] + 1:
                run[1] = i
                continue
            _lib.check(lib.sg2_counter_add(self._present.data_ptr(), run[0], run[1] - run[0] + 1, 1, 0, st), 'sg2_counter_add')
            run = [i, i]
        _lib.check(lib.sg2_adam_multi(
            self._P.data_ptr(), self._G.data_ptr(), self._M.data_ptr(), self._V.data_ptr(), self._total,
            self._seg_off.data_ptr(), self._steps.data_ptr(), self._present.data_ptr(), self._coef.data_ptr(), len(self._ps),
            float(lr), float(b1), float(b2), float(eps), float(scale), st), 'sg2_adam_multi')
