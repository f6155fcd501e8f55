"""Training utilities of the reference (nnutils/__init__.py) on the B200 path."""
from __future__ import annotations

import os

import torch

from .accelerate import MiniAccelerator, init_distributed            # noqa: F401
from .optim import FlatAdam                                           # noqa: F401
from .training import sample_nnoise, sample_unoise, update_ema        # noqa: F401
from . import loss                                                    # noqa: F401


def get_device(gpu: bool = True):
    """cuda:LOCAL_RANK (cuda:0 outside torchrun, as the reference nnutils/__init__.py:18-21) or cpu."""
    if gpu and torch.cuda.is_available():
        return torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    return torch.device('cpu')


def freeze(model: torch.nn.Module) -> None:
    model.eval()
    for p in model.parameters():
        p.requires_grad = False


def unfreeze(model: torch.nn.Module) -> None:
    for p in model.parameters():
        p.requires_grad = True
    model.train()
