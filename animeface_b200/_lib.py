"""ctypes binding of libsg2b200.so (the C ABI declared in include/sg2b200.h).

This is the only place the shared library is opened.  There is NO fallback: if the library is
missing or fails to load, every op raises -- the product path never routes through PyTorch
composites or the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsg2b200.so')

_lib = None

_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p
_int = C.c_int
_f = C.c_float
_i64 = C.c_int64

# name -> (restype, argtypes); mirrors include/sg2b200.h one to one.
SIGNATURES = {
    'sg2_version': (_int, []),
    'sg2_last_error': (C.c_char_p, []),
    'sg2_launch_count': (_i64, []),
    'sg2_debug_trace': (_int, [_vp]),
    'sg2_upfirdn2d': (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _i64p, _int, _int, _i64p,
                             _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _f, _vp]),
    'sg2_up2x_fwd': (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _vp]),
    'sg2_up2x_adj': (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _vp]),
    'sg2_avgpool2_fwd': (_int, [_vp, _vp, _vp, _vp, _f, _int, _int, _int, _int, _int, _vp]),
    'sg2_avgpool2_adj': (_int, [_vp, _vp, _f, _int, _int, _int, _int, _vp]),
    'sg2_bias_act': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _i64, _int, _i64, _int, _int, _f, _f, _f, _vp]),
    'sg2_mbstd_fwd': (_int, [_vp, _i64p, _vp, _i64p, _vp, _int, _int, _int, _int, _int, _f, _vp]),
    'sg2_mbstd_bwd': (_int, [_vp, _i64p, _vp, _i64p, _vp, _i64p, _int, _int, _int, _int, _int, _f, _vp]),
    'sg2_conv2d_select_impl': (_int, [_int, _int, _int, _int, _int, _int, _int, _int]),
    'sg2_conv2d_packed_size': (_i64, [_int, _int, _int, _int]),
    'sg2_conv2d_pack_weight': (_int, [_vp, _vp, _int, _int, _int, _f, _int, _int, _vp]),
    'sg2_conv2d_fwd': (_int, [_vp, _vp, _vp, _i64p, _int, _int, _int, _int, _int, _int,
                              _vp, _vp, _vp, _vp, _int, _f, _f, _int, _vp]),
    'sg2_conv2d_wgrad_workspace': (_i64, [_int, _int, _int, _int, _int, _int, _int]),
    'sg2_conv2d_wgrad': (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _f, _vp, _vp, _int, _int, _vp, _vp]),
    'sg2_reduce_hw_workspace': (_i64, [_int, _int, _int]),
    'sg2_reduce_hw': (_int, [_vp, _vp, _vp, _int, _int, _int, _vp, _vp]),
    'sg2_scale_reduce_hw': (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _vp, _vp]),
    'sg2_modconv_bwd_prep': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _f, _vp, _vp]),
    'sg2_split_planes': (_int, [_vp, _vp, _vp, _int, _int, _int, _vp]),
    'sg2_bwd_prep_planes_workspace': (_i64, [_int, _int, _int]),
    'sg2_bwd_prep_planes': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _f, _int, _f, _vp]),
    'sg2_conv2d_planes_supported': (_int, [_int, _int, _int, _int, _int, _int, _int]),
    'sg2_conv2d_fwd_planes': (_int, [_vp, _vp, _vp, _i64p, _int, _int, _int, _int, _int, _int, _vp, _vp, _int, _f, _f, _int, _vp]),
    'sg2_conv2d_wgrad_planes_workspace': (_i64, [_int, _int, _int, _int, _int, _int]),
    'sg2_conv2d_wgrad_planes': (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _f, _int, _vp]),
    'sg2_thin_in_bwd_workspace': (_i64, [_int, _int, _int, _int]),
    'sg2_thin_in_bwd': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _f, _f, _f, _vp]),
    'sg2_demod_fwd': (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _f, _f, _vp]),
    'sg2_demod_bwd': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _f, _vp]),
    'sg2_filtered_lrelu': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int] + [_int] * 16 + [_f, _f, _f, _f, _int, _vp]),
    'sg2_linear_fwd_workspace': (_i64, [_int, _int, _int]),
    'sg2_linear_fwd': (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _f, _f, _f, _vp, _vp]),
    'sg2_linear_bwd_data': (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _f, _f, _f, _vp]),
    'sg2_linear_bwd_weight': (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _f, _f, _f, _vp]),
    'sg2_pixelnorm': (_int, [_vp, _vp, _int, _int, _f, _vp]),
    'sg2_diffaugment_workspace': (_i64, [_int, _int, _int]),
    'sg2_diffaugment': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _vp, _vp]),
    'sg2_reflect_pad': (_int, [_vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _int, _vp]),
    'sg2_affine_sample': (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _vp]),
    'sg2_color_affine': (_int, [_vp, _vp, _vp, _int, _i64, _int, _vp]),
    'sg2_ema_update': (_int, [_vp, _vp, _i64, _f, _vp]),
    'sg2_counter_add': (_int, [_vp, _int, _int, _int, _int, _vp]),
    'sg2_adam_multi': (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _int, _f, _f, _f, _f, _f, _vp]),
    'sg2_adam_ema': (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _f, _f, _vp]),
}


def load():
    """Open the library (once) and declare every prototype.  Raises if it cannot be opened."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or `make -C animeface_b200/csrc`). There is no fallback path.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: loud by design
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = load().sg2_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{what or "libsg2b200"} failed ({rc}): {msg}')


def stream_ptr(t: torch.Tensor | None = None) -> int:
    return torch.cuda.current_stream(t.device if t is not None else None).cuda_stream


def ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def strides4(t: torch.Tensor):
    assert t.ndim == 4
    return (C.c_int64 * 4)(*t.stride())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('libsg2b200 ops need CUDA tensors (there is no CPU path)')


DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.float64: 2}


# Autocast (the reference's default run mode is fp16 autocast + GradScaler: implementations/StyleGAN2/utils.py:47,62,167,
# utils/argument.py:25): every custom autograd Function of this package takes its tensors as float32 -- inside an
# autocast region incoming half tensors are cast up and the kernels compute and store fp32 ("AMP-compatible, fp32 compute").
amp_fwd = torch.amp.custom_fwd(device_type='cuda', cast_inputs=torch.float32)
amp_bwd = torch.amp.custom_bwd(device_type='cuda')


def launch_count() -> int:
    return int(load().sg2_launch_count())
