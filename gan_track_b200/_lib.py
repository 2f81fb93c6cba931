"""ctypes binding of the C ABI declared in `include/gantrack_b200.h`.

There is NO fallback: if `libgantrack_b200.so` is missing or an entry point reports an error, the caller gets an
exception.  The library links the shared libcudart, so torch is imported first (it loads libcudart.so.12).
"""
import ctypes
import os
import threading

import torch  # noqa: F401  (loads libcudart before our library is opened)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgantrack_b200.so')

_c = ctypes
_vp, _i, _ll, _f = _c.c_void_p, _c.c_int, _c.c_longlong, _c.c_float

# name -> (restype, argtypes); one line per prototype of include/gantrack_b200.h
_PROTOTYPES = {
    'gt_last_error': (_c.c_char_p, []),
    'gt_abi_version': (_i, []),
    'gt_sm_count': (_i, []),
    'gt_stream_config': (_i, [_i]),
    'gt_bias_act': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _ll, _i, _ll, _vp]),
    'gt_bias_act_bwd_workspace': (_ll, [_i, _i, _ll]),
    'gt_bias_act_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _f, _f, _i, _i, _ll, _vp]),
    'gt_upfirdn2d': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _i, _ll, _ll, _i, _i, _ll, _ll, _ll, _ll,
                          _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    'gt_aug_params': (_i, [_c.POINTER(_c.c_void_p), _vp, _c.POINTER(_c.c_float), _c.POINTER(_c.c_float), _i, _i, _i, _i, _vp, _vp, _vp]),
    'gt_aug_warp_workspace': (_ll, [_i, _i, _i, _i]),
    'gt_aug_warp_fwd': (_i, [_vp, _vp, _vp, _c.POINTER(_c.c_float), _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _ll, _vp]),
    'gt_aug_warp_bwd': (_i, [_vp, _vp, _vp, _c.POINTER(_c.c_float), _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _ll, _vp]),
    'gt_mod_workspace': (_ll, [_i, _ll, _i, _i]),
    'gt_mod_scale_fwd': (_i, [_vp, _vp, _vp, _i, _i, _ll, _i, _vp]),
    'gt_mod_scale_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _i, _vp]),
    'gt_demod_act_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _i, _ll, _i, _vp]),
    'gt_demod_act_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _f, _f, _i, _ll, _i, _vp]),
    'gt_mod_scale_bwd2': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _i, _vp]),
    'gt_demod_act_bwd2': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _f, _f, _i, _ll, _i, _vp]),
    'gt_fc_fwd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _vp]),
    'gt_fc_dgrad': (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    'gt_fc_wgrad': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _vp]),
    'gt_batch_gather': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _vp]),
    'gt_modprep_weight_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'gt_modprep_style_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'gt_modprep_rsqrt': (_i, [_vp, _vp, _i, _f, _vp]),
    'gt_modprep_gq': (_i, [_vp, _vp, _vp, _i, _vp]),
    'gt_modprep_style_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'gt_modprep_weight_bwd': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'gt_modprep_style_bwd2_a': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'gt_modprep_style_bwd2_b': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'gt_modprep_style_bwd2_c': (_i, [_vp] * 12 + [_i, _i, _i, _vp]),
    'gt_adam_chunk_bytes': (_i, []),
    'gt_adam_flat': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _f, _f, _f, _f, _f, _f, _f, _vp]),
    'gt_ema_flat': (_i, [_vp, _vp, _ll, _f, _vp]),
    'gt_fromrgb1_bwd_workspace': (_ll, [_i]),
    'gt_fromrgb1_fwd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _ll, _i, _vp]),
    'gt_fromrgb1_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _f, _f, _ll, _i, _vp]),
    'gt_torgb1_bwd_workspace': (_ll, [_i, _i]),
    'gt_torgb1_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _f, _i, _ll, _i, _vp]),
    'gt_torgb1_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _i, _vp]),
    'gt_conv_igemm_config': (_i, [_i]),
    'gt_conv_rows_config': (_i, [_i]),
    'gt_conv_wgrad_config': (_i, [_i]),
    'gt_conv_pack_weight_f16': (_i, [_vp, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _vp, _vp]),
    'gt_conv2d_igemm_f16': (_i, [_vp, _ll, _ll, _ll, _vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'gt_f16x3_amax': (_i, [_vp, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _vp, _vp]),
    'gt_f16x3_split_act': (_i, [_vp, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'gt_f16x3_pack_weight': (_i, [_vp, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'gt_conv2d_igemm_f16_f32out': (_i, [_vp, _ll, _ll, _ll, _vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _ll, _vp]),
    'gt_f16x3_slab_reduce': (_i, [_vp, _ll, _i, _ll, _i, _i, _vp, _vp, _vp, _vp]),
    'gt_conv2d_wgrad_f16x3_workspace': (_ll, [_i, _i, _i, _i, _i, _i, _i]),
    'gt_conv2d_wgrad_f16x3': (_i, [_vp, _ll, _ll, _ll, _i, _i, _i, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _ll, _ll, _ll, _ll, _i, _i, _vp, _vp, _vp, _ll, _vp]),
    'gt_conv2d_igemm_f16_bias_act': (_i, [_vp, _ll, _ll, _ll, _vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _f, _f, _f, _vp]),
    'gt_conv2d_igemm_f16_bias_act_add': (_i, [_vp, _ll, _ll, _ll, _vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _f, _f, _f, _vp, _vp]),
    'gt_conv2d_wgrad_workspace': (_ll, [_i, _i, _i, _i, _i, _i, _i]),
    'gt_conv2d_wgrad_f16': (_i, [_vp, _ll, _ll, _ll, _i, _i, _i, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _ll, _ll, _ll, _ll, _vp, _ll, _vp]),
}

_lib = None
_lock = threading.Lock()

# Number of kernel launches issued through the C ABI by this process (bench.py reports it as `gpu_launches`).
launches = 0


def count_launch(n=1):
    global launches
    launches += n


class GanTrackLibraryError(RuntimeError):
    pass


def load():
    """Open the shared library (once) and attach prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GanTrackLibraryError(
                f'{LIB_PATH} not found: build it with `python -m gan_track_b200.build` (there is no CPU or library fallback)')
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def is_built():
    return os.path.exists(LIB_PATH)


def check(status, what):
    if status != 0:
        msg = load().gt_last_error()
        raise RuntimeError(f'{what} failed (status {status}): {msg.decode() if msg else "unknown error"}')


DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.float64: 2}


def dtype_code(t):
    try:
        return DTYPE_CODE[t.dtype]
    except KeyError:
        raise RuntimeError(f'gan_track_b200: unsupported dtype {t.dtype}') from None


def ptr(t):
    """Device pointer of a tensor, or None (NULL) for None / empty tensors."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f'gan_track_b200: {name} must be a CUDA tensor (this package has no CPU path; device={t.device})')
