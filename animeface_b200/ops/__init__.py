"""Operators of the reference's StyleGAN paths, each a thin autograd wrapper over libsg2b200 entry points.
Module names follow thirdparty/stylegan3_ops/ops/ (upfirdn2d, bias_act, conv2d_gradfix, conv2d_resample, filtered_lrelu,
grid_sample -> grid_sample_gradfix); conv2d / linear / mbstd / resample carry the fused forms the networks here call."""
from . import (bias_act, conv2d, conv2d_gradfix, conv2d_resample, filtered_lrelu, grid_sample, linear, mbstd, resample,  # noqa: F401
               upfirdn2d)
