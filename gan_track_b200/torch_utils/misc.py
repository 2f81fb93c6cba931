"""Host helpers used by the networks (counterparts of S3/torch_utils/misc.py:20-105, 145-175)."""
import contextlib

import numpy as np
import torch

_constant_cache = dict()


def constant(value, shape=None, dtype=None, device=None, memory_format=None):
    """Cached constant tensor (avoids a host->device copy per use)."""
    value = np.asarray(value)
    shape = tuple(shape) if shape is not None else None
    dtype = dtype or torch.get_default_dtype()
    device = torch.device(device) if device is not None else torch.device('cpu')
    memory_format = memory_format or torch.contiguous_format
    key = (value.shape, value.dtype, value.tobytes(), shape, dtype, device, memory_format)
    t = _constant_cache.get(key)
    if t is None:
        t = torch.as_tensor(value.copy(), dtype=dtype, device=device)
        if shape is not None:
            t = t.expand(shape)
        t = t.contiguous(memory_format=memory_format)
        _constant_cache[key] = t
    return t


nan_to_num = torch.nan_to_num


def assert_shape(tensor, ref_shape):
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f'Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}')
    for idx, (size, ref) in enumerate(zip(tensor.shape, ref_shape)):
        if ref is not None and int(size) != int(ref):
            raise AssertionError(f'Wrong size for dimension {idx}: got {size}, expected {ref}')


def profiled_function(fn):
    def wrapper(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    wrapper.__name__ = fn.__name__
    wrapper.__doc__ = fn.__doc__
    return wrapper


@contextlib.contextmanager
def suppress_tracer_warnings():
    yield


def params_and_buffers(module):
    return list(module.parameters()) + list(module.buffers())


def named_params_and_buffers(module):
    return list(module.named_parameters()) + list(module.named_buffers())


def copy_params_and_buffers(src_module, dst_module, require_all=False):
    src = dict(named_params_and_buffers(src_module))
    for name, tensor in named_params_and_buffers(dst_module):
        assert (name in src) or (not require_all), name
        if name in src:
            tensor.copy_(src[name].detach()).requires_grad_(tensor.requires_grad)


def check_ddp_consistency(module, ignore_regex=None):
    """Replicas must hold bit-identical parameters (S3/torch_utils/misc.py:180-191)."""
    import re
    for name, tensor in named_params_and_buffers(module):
        fullname = type(module).__name__ + '.' + name
        if ignore_regex is not None and re.fullmatch(ignore_regex, fullname):
            continue
        tensor = tensor.detach()
        if tensor.is_floating_point():
            tensor = nan_to_num(tensor)
        other = tensor.clone()
        torch.distributed.broadcast(tensor=other, src=0)
        assert (tensor == other).all(), fullname
