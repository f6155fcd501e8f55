"""GAN losses and gradient penalties with the reference's names.

  NonSaturatingLoss / Adversarial   nnutils/loss/gan.py:8-38, 98-114
  calc_grad                         nnutils/loss/penalty.py:11-26
  r1_regularizer / r2_regularizer   nnutils/loss/penalty.py:85-108
  gradient_penalty                  nnutils/loss/penalty.py:33-58
The penalties rely on second-order autograd through the discriminator; every op of
animeface_b200.model.Discriminator is twice differentiable on the library kernels.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F


class Loss:
    def __init__(self, return_all: bool = False) -> None:
        self.return_all = return_all


class Adversarial(Loss):
    def real_loss(self, prob):
        raise NotImplementedError()

    def fake_loss(self, prob):
        raise NotImplementedError()

    def d_loss(self, real_prob, fake_prob):
        rl, fl = self.real_loss(real_prob), self.fake_loss(fake_prob)
        return (rl + fl, rl, fl) if self.return_all else rl + fl

    def g_loss(self, fake_prob):
        return self.real_loss(fake_prob)


class NonSaturatingLoss(Adversarial):
    def real_loss(self, prob):
        return F.softplus(-prob).mean()

    def fake_loss(self, prob):
        return F.softplus(prob).mean()


class HingeLoss(Adversarial):
    def real_loss(self, prob):
        return F.relu(1. - prob).mean()

    def fake_loss(self, prob):
        return F.relu(1. + prob).mean()

    def g_loss(self, fake_prob):
        return -fake_prob.mean()


class WGANLoss(Adversarial):
    def real_loss(self, prob):
        return -prob.mean()

    def fake_loss(self, prob):
        return prob.mean()


def _is_scaler(s):
    return s is not None and hasattr(s, 'scale') and hasattr(s, 'get_scale')


def calc_grad(outputs, inputs, scaler=None):
    """d(sum outputs)/d inputs with create_graph (reference nnutils/loss/penalty.py:11-26)."""
    with torch.autocast('cuda', enabled=False):
        if _is_scaler(scaler):
            outputs = scaler.scale(outputs)
        ones = torch.ones(outputs.size(), device=outputs.device)
        # only the gradient of `inputs` (an image or a latent) is wanted: the convolution nodes skip their weight gradients, as
        # ATen's own convolution does for outputs nobody asked for (ops.conv2d.skip_weight_grads)
        from ..ops.conv2d import skip_weight_grads
        wanted = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        skip = not any(isinstance(t, torch.nn.Parameter) for t in wanted)
        with (skip_weight_grads() if skip else contextlib.nullcontext()):
            gradients = torch.autograd.grad(outputs=outputs, inputs=inputs, grad_outputs=ones,
                                            create_graph=True, retain_graph=True, only_inputs=True)[0]
        if _is_scaler(scaler):
            gradients = gradients / scaler.get_scale()
    return gradients


class Penalty(Loss):
    def __init__(self, return_all: bool = False) -> None:
        super().__init__(return_all=return_all)
        self.filter_output = lambda x: x


class r1_regularizer(Penalty):
    """mean_b(||d D(x)/dx||^2) / 2 on real samples (reference nnutils/loss/penalty.py:85-101)."""

    def __call__(self, real, D, scaler=None, d_aux_input=tuple()):
        real_loc = real.detach().requires_grad_(True)
        d_real_loc = self.filter_output(D(real_loc, *d_aux_input))
        gradients = calc_grad(d_real_loc, real_loc, scaler)
        gradients = gradients.reshape(gradients.size(0), -1)
        return gradients.norm(2, dim=1).pow(2).mean() / 2.


class r2_regularizer(r1_regularizer):
    def __call__(self, fake, D, scaler=None, d_aux_input=tuple()):
        return super().__call__(fake, D, scaler, d_aux_input)


class gradient_penalty(Penalty):
    def __call__(self, real, fake, D, scaler=None, center=1., d_aux_input=tuple()):
        assert center in [1., 0.]
        alpha = torch.rand(1, device=real.device)
        x_hat = (real * alpha + fake * (1 - alpha)).detach().requires_grad_(True)
        d_x_hat = self.filter_output(D(x_hat, *d_aux_input))
        gradients = calc_grad(d_x_hat, x_hat, scaler).reshape(real.size(0), -1)
        return (gradients.norm(2, dim=1) - center).pow(2).mean()
