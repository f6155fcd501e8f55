"""FlatAdam: torch.optim.Adam semantics over one flat fp32 buffer, one kernel, one all-reduce.

Replaces ``optim.Adam(G.parameters(), ...)`` / ``optim.Adam(D.parameters(), ...)`` of
implementations/StyleGAN2/utils.py:208-221 (and, with ``ema_model=``, the 81 per-tensor lerps of
nnutils/training.py:23-40 become one kernel in ``update_ema``).

  * parameters are re-pointed to views of ONE flat buffer; m / v are flat too;
  * ``step()`` packs the gradients with one multi-tensor copy, all-reduces the flat buffer ONCE when
    torch.distributed is initialised (the only collective of the data-parallel step; there is no DDP in the
    reference, nnutils/accelerate.py:8-11), and launches ``sg2_adam_ema`` per run of tensors;
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
        self._host_steps = [0] * len(ps)
        for p, o, n in zip(ps, self._offs, self._sizes):
            self._P[o:o + n].copy_(p.data.reshape(-1))
            p.data = self._P[o:o + n].view(p.shape)
        self._gviews = [self._G[o:o + n].view(p.shape) for p, o, n in zip(ps, self._offs, self._sizes)]
        self._idx_cache = {}
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

    def zero_grad(self, set_to_none: bool = True):
        for p in self._ps:
            p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        group = self.param_groups[0]
        lr, (b1, b2), eps = group['lr'], group['betas'], group['eps']
        present = [i for i, p in enumerate(self._ps) if p.grad is not None]
        if not present:
            return
        torch._foreach_copy_([self._gviews[i] for i in present], [self._ps[i].grad for i in present])
        scale = 1.0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self._G)                      # the one collective of the step
            scale = 1.0 / dist.get_world_size()
        key = tuple(present)
        if key not in self._idx_cache:
            self._idx_cache[key] = torch.tensor(present, dtype=torch.int64, device=self._P.device)
        self._steps.index_add_(0, self._idx_cache[key], torch.ones_like(self._idx_cache[key]))
        for i in present:
            self._host_steps[i] += 1
        lib = _lib.load()
        st = _lib.stream_ptr(self._P)
        # maximal runs of consecutive present tensors sharing a step count -> one launch each
        run = [present[0]]
        runs = []
        for i in present[1:]:
            if i == run[-1] + 1 and self._host_steps[i] == self._host_steps[run[0]]:
                run.append(i)
            else:
                runs.append(run)
                run = [i]
        runs.append(run)
        for r in runs:
            o0 = self._offs[r[0]]
            o1 = self._offs[r[-1]] + self._sizes[r[-1]]
            _lib.check(lib.sg2_adam_ema(
                self._P.data_ptr() + 4 * o0, self._G.data_ptr() + 4 * o0, self._M.data_ptr() + 4 * o0,
                self._V.data_ptr() + 4 * o0, None, o1 - o0, self._steps.data_ptr() + 8 * r[0],
                float(lr), float(b1), float(b2), float(eps), float(scale), 0.0, st), 'sg2_adam_ema')
