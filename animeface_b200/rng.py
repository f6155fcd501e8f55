"""Random draws of the training step, in one place so tests can replay the reference's draws.

The reference takes every random number from the global torch generator in a fixed call order
(implementations/StyleGAN2/utils.py:61,64,67-68,89,92-93; model.py:86-88; DiffAugment.py:24,30,36,42-43).
The product path does the same on the device; under ``replay(seq)`` the draws are popped from ``seq``
instead, which is how CPU-reference trajectories are reproduced on the GPU.
"""
from __future__ import annotations

import torch

_queue = None


class replay:
    def __init__(self, seq):
        self.seq = list(seq)

    def __enter__(self):
        global _queue
        self._prev = _queue
        _queue = self.seq
        return self

    def __exit__(self, *exc):
        global _queue
        _queue = self._prev

    @property
    def remaining(self):
        return len(self.seq)


def _pop(shape, device):
    t = _queue.pop(0)
    assert tuple(t.shape) == tuple(shape), f'replayed draw has shape {tuple(t.shape)}, expected {tuple(shape)}'
    return t.to(device)


def randn(*shape, device):
    return _pop(shape, device) if _queue is not None else torch.randn(*shape, device=device)


def rand(*shape, device, dtype=torch.float32):
    return _pop(shape, device).to(dtype) if _queue is not None else torch.rand(*shape, device=device, dtype=dtype)


def randint(lo, hi, shape, device):
    return _pop(shape, device) if _queue is not None else torch.randint(lo, hi, size=list(shape), device=device)
