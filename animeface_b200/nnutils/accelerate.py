"""MiniAccelerator with the reference's API (nnutils/accelerate.py:134-252) plus the data-parallel layer
the reference does not have (its header: "NOT implemented: Multi-device", accelerate.py:8-11).

Same surface: ``prepare(*args)``, ``backward(loss)``, ``autocast()``, ``update()``, ``.scaler``, ``.device``.
Added behaviour when launched under torchrun (one process per GPU, NCCL over NVLink):
  * ``device`` defaults to ``cuda:LOCAL_RANK`` and the process group is initialised on first use;
  * ``prepare(model)`` broadcasts rank 0's parameters so replicas start identical;
  * ``prepare(optimizer)`` wraps a plain torch optimizer so that ``step()`` first all-reduces (mean) ONE flat
    gradient buffer -- a ``FlatAdam`` does that itself inside its fused step;
  * ``prepare(dataloader)`` shards by rank when the loader has a plain sampler.
There is exactly one collective per optimizer step; nothing else is exchanged.
"""
from __future__ import annotations

import os
from contextlib import contextmanager

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader

from .optim import FlatAdam


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def init_distributed(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    if not dist.is_available() or dist.is_initialized():
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group(backend=backend, rank=int(os.environ.get('RANK', '0')), world_size=world)


class AllReduceOptimizer(optim.Optimizer):
    """Wraps any torch optimizer: mean-all-reduce of one flat gradient buffer, then the wrapped step.
    Parameters without a gradient contribute zeros so every rank reduces the same layout."""

    def __init__(self, optimizer: optim.Optimizer, scaler=None):
        self._optimizer = optimizer
        self._scaler = scaler
        self._params = [p for g in optimizer.param_groups for p in g['params']]
        self._flat = None

    def _allreduce(self):
        world = _world()
        if world <= 1:
            return
        if self._flat is None:
            n = sum(p.numel() for p in self._params)
            self._flat = torch.zeros(n, dtype=torch.float32, device=self._params[0].device)
        flat, off = self._flat, 0
        flat.zero_()
        views = []
        for p in self._params:
            v = flat[off:off + p.numel()].view(p.shape)
            views.append(v)
            off += p.numel()
        have = [(v, p.grad) for v, p in zip(views, self._params) if p.grad is not None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        dist.all_reduce(flat)
        flat.mul_(1.0 / world)
        if have:
            torch._foreach_copy_([g for _, g in have], [v for v, _ in have])

    def step(self, closure=None):
        self._allreduce()
        if self._scaler is not None:
            self._scaler.step(self._optimizer, closure)
        else:
            self._optimizer.step(closure)

    def zero_grad(self, set_to_none=None):
        self._optimizer.zero_grad(set_to_none=False if set_to_none is None else set_to_none)

    @property
    def param_groups(self):
        return self._optimizer.param_groups

    @param_groups.setter
    def param_groups(self, v):
        self._optimizer.param_groups = v

    @property
    def defaults(self):
        return self._optimizer.defaults

    def add_param_group(self, g):
        self._optimizer.add_param_group(g)

    def state_dict(self):
        return self._optimizer.state_dict()

    def load_state_dict(self, sd):
        self._optimizer.load_state_dict(sd)


class ScaledFlatAdam:
    """FlatAdam behind a GradScaler (amp=True): exchange the scaled gradients first, then let the scaler unscale them,
    check them for infs (every rank sees the same reduced values, so every rank takes the same decision) and step."""

    def __init__(self, optimizer: FlatAdam, scaler):
        self._optimizer, self._scaler = optimizer, scaler

    def step(self, closure=None):
        assert closure is None
        self._optimizer.reduce_gradients()
        self._scaler.step(self._optimizer)

    def zero_grad(self, set_to_none=True):
        self._optimizer.zero_grad(set_to_none)

    def __getattr__(self, name):
        return getattr(self._optimizer, name)


class DataLoaderWrapper(DataLoader):
    """Moves every batch to the device (reference accelerate.py:98-132)."""

    def __init__(self, dataset, device, **kwargs):
        super().__init__(dataset, **kwargs)
        self._device = device
        self._epoch = 0

    def __iter__(self):
        if isinstance(self.sampler, torch.utils.data.distributed.DistributedSampler):
            self.sampler.set_epoch(self._epoch)          # a new shuffle per epoch, the same one on every rank
        self._epoch += 1
        for batch in super().__iter__():
            yield _to_device(batch, self._device)

    @classmethod
    def from_dataloader(cls, dataloader: DataLoader, device):
        sampler_kwargs = dict(batch_sampler=dataloader.batch_sampler)
        world = _world()
        if world > 1 and isinstance(dataloader.sampler, (torch.utils.data.RandomSampler, torch.utils.data.SequentialSampler)):
            shuffle = isinstance(dataloader.sampler, torch.utils.data.RandomSampler)
            sampler = torch.utils.data.distributed.DistributedSampler(dataloader.dataset, shuffle=shuffle)
            sampler_kwargs = dict(batch_size=dataloader.batch_size, sampler=sampler, drop_last=dataloader.drop_last)
        elif world > 1 and not isinstance(dataloader.sampler, torch.utils.data.distributed.DistributedSampler):
            raise ValueError('MiniAccelerator.prepare: with world_size > 1 a DataLoader needs a plain Random/Sequential sampler '
                             '(it is re-created with a DistributedSampler) or its own DistributedSampler; a custom sampler / '
                             'batch_sampler cannot be sharded by rank automatically')
        return cls(dataloader.dataset, device, num_workers=dataloader.num_workers,
                   pin_memory=dataloader.pin_memory, **sampler_kwargs)


def _to_device(data, device):
    if isinstance(data, (tuple, list)):
        return type(data)(_to_device(e, device) for e in data)
    if isinstance(data, dict):
        return {k: _to_device(v, device) for k, v in data.items()}
    if isinstance(data, torch.Tensor) and device is not None:
        return data.to(device, non_blocking=True)
    return data


class MiniAccelerator:
    def __init__(self, amp: bool = True, device_placement: bool = True, device=None) -> None:
        init_distributed()
        self._amp = amp
        self._device_placement = device_placement
        if device is None:
            if torch.cuda.is_available():
                device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
            else:
                device = torch.device('cpu')
        self._device = device
        self._scaler = torch.amp.GradScaler('cuda') if amp else None

    @property
    def scaler(self):
        return self._scaler

    @property
    def device(self):
        return self._device

    @device.setter
    def device(self, device):
        self._device = device

    @property
    def world_size(self):
        return _world()

    def update(self):
        if self._scaler is not None:
            self._scaler.update()

    def backward(self, loss: torch.Tensor):
        if self._scaler is not None:
            self._scaler.scale(loss).backward()
        else:
            loss.backward()

    def prepare(self, *args):
        if len(args) == 0:
            return
        out = []
        for a in args:
            if isinstance(a, nn.Module):
                out.append(self._prepare_model(a))
            elif isinstance(a, optim.Optimizer):
                out.append(self._prepare_optimizer(a))
            elif isinstance(a, DataLoader):
                out.append(self._prepare_dataloader(a))
            else:
                out.append(a)
        return tuple(out) if len(out) > 1 else out[0]

    def _prepare_model(self, model: nn.Module) -> nn.Module:
        if self._device_placement:
            model.to(self._device)
        if _world() > 1:
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, src=0)
        return model

    def _prepare_optimizer(self, optimizer):
        if isinstance(optimizer, FlatAdam):
            # all-reduce is fused into its step; behind a GradScaler the exchange has to come before the scaler's inf check
            return ScaledFlatAdam(optimizer, self._scaler) if self._scaler is not None else optimizer
        if self._scaler is not None or _world() > 1:
            return AllReduceOptimizer(optimizer, self._scaler)
        return optimizer

    def _prepare_dataloader(self, dataloader):
        if self._device_placement:
            dataloader = DataLoaderWrapper.from_dataloader(dataloader, self._device)
        return dataloader

    @contextmanager
    def autocast(self):
        if self._amp:
            with torch.autocast(self._device.type if isinstance(self._device, torch.device) else 'cuda'):
                yield
        else:
            yield
