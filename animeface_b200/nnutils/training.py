"""Training utilities with the reference's names (nnutils/training.py:7-40)."""
from __future__ import annotations

import torch

from .. import _lib


def sample_nnoise(size, device, mean: float = 0., std: float = 1.) -> torch.Tensor:
    return torch.empty(size, device=device).normal_(mean, std)


def sample_unoise(size, device, start: float = 0., end: float = 1.) -> torch.Tensor:
    return torch.empty(size, device=device).uniform_(start, end)


@torch.no_grad()
def update_ema(model, model_ema, decay: float = 0.999, copy_buffers: bool = False) -> None:
    """p_ema = decay * p_ema + (1 - decay) * p over named_parameters (reference nnutils/training.py:23-40).

    When both models were flattened by ``FlatAdam`` (``ema_model=`` argument) this is ONE kernel over the
    two flat fp32 buffers instead of one lerp per tensor (81 for the StyleGAN2 generator)."""
    model.eval()
    flat, flat_ema = getattr(model, '_sg2_flat', None), getattr(model_ema, '_sg2_flat', None)
    if flat is not None and flat_ema is not None and flat.numel() == flat_ema.numel() and flat.is_cuda:
        _lib.check(_lib.load().sg2_ema_update(flat_ema.data_ptr(), flat.data_ptr(), flat.numel(), float(decay),
                                              _lib.stream_ptr(flat)), 'sg2_ema_update')
    else:
        src = dict(model.named_parameters())
        for key, p_ema in model_ema.named_parameters():
            p_ema.data.mul_(decay).add_(src[key].data, alpha=(1 - decay))
    if copy_buffers:
        src = dict(model.named_buffers())
        for key, b_ema in model_ema.named_buffers():
            b_ema.data.copy_(src[key].data)
    model.train()
