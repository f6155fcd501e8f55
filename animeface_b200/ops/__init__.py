"""Operators of the StyleGAN2 hot path, each a thin autograd wrapper over one libsg2b200 entry point."""
from . import bias_act, conv2d, conv2d_gradfix, conv2d_resample, mbstd, resample, upfirdn2d  # noqa: F401
