"""Plugin loader with the reference's signature: `get_plugin(module_name, sources, headers, source_dir, **kw)`.

The reference JIT-compiles a pybind module with ninja and returns it (S3/torch_utils/custom_ops.py:59-155).  Here
nothing is compiled at import: the kernels are prebuilt into `libgantrack_b200.so` (C ABI, include/gantrack_b200.h)
and this function hands back a thin object with the SAME methods the reference's op wrappers call:

    get_plugin('bias_act_plugin', ...).bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)
    get_plugin('upfirdn2d_plugin', ...).upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)

so the reference's unmodified `OPS/bias_act.py` / `OPS/upfirdn2d.py` run on top of it.  Outputs are
allocated with torch (memory stays in the caching allocator, as with `torch::empty_like` in OPS/bias_act.cpp:55).
"""
import torch

from .. import _lib

verbosity = 'none'  # kept for interface compatibility ('none' | 'brief' | 'full')


class _BiasActPlugin:
    """Counterpart of the pybind module built from OPS/bias_act.cpp:94-97."""

    @staticmethod
    def bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp):
        _lib.require_cuda(x, 'x')
        for name, t in (('b', b), ('xref', xref), ('yref', yref), ('dy', dy)):
            if t.numel() and (t.dtype != x.dtype or t.device != x.device):
                raise RuntimeError(f'{name} must have the same dtype and device as x')
        for name, t in (('xref', xref), ('yref', yref), ('dy', dy)):
            if t.numel() and (t.shape != x.shape or t.stride() != x.stride()):
                raise RuntimeError(f'{name} must have the same shape and layout as x')
        if b.numel():
            if b.ndim != 1:
                raise RuntimeError('b must have rank 1')
            if not 0 <= dim < x.ndim:
                raise RuntimeError('dim is out of bounds')
            if b.numel() != x.shape[dim]:
                raise RuntimeError('b has wrong number of elements')
            if not b.is_contiguous():
                raise RuntimeError('b must be contiguous')
        if grad < 0:
            raise RuntimeError('grad must be non-negative')
        if not _is_dense(x):
            raise RuntimeError('x must be non-overlapping and dense')
        y = torch.empty_like(x)
        if x.numel() == 0:
            return y
        step_b = x.stride(dim) if b.numel() else 1
        with torch.cuda.device(x.device):
            st = _lib.load().gt_bias_act(_lib.ptr(x), _lib.ptr(b), _lib.ptr(xref), _lib.ptr(yref), _lib.ptr(dy), _lib.ptr(y),
                                         _lib.dtype_code(x), int(grad), int(act), float(alpha), float(gain), float(clamp),
                                         x.numel(), int(b.numel()), int(step_b), _lib.stream_of(x))
        _lib.check(st, 'bias_act')
        _lib.count_launch()
        return y


class _Upfirdn2dPlugin:
    """Counterpart of the pybind module built from OPS/upfirdn2d.cpp:102-105."""

    @staticmethod
    def upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        _lib.require_cuda(x, 'x')
        if f.device != x.device:
            raise RuntimeError('f must reside on the same device as x')
        if f.dtype != torch.float32:
            raise RuntimeError('f must be float32')
        if x.ndim != 4:
            raise RuntimeError('x must be rank 4')
        if f.ndim != 2:
            raise RuntimeError('f must be rank 2')
        if x.numel() == 0:
            raise RuntimeError('x has zero size')
        if f.numel() == 0:
            raise RuntimeError('f has zero size')
        if upx < 1 or upy < 1:
            raise RuntimeError('upsampling factor must be at least 1')
        if downx < 1 or downy < 1:
            raise RuntimeError('downsampling factor must be at least 1')
        n, c, h, w = x.shape
        fh, fw = f.shape
        ow = (w * upx + padx0 + padx1 - fw + downx) // downx
        oh = (h * upy + pady0 + pady1 - fh + downy) // downy
        if ow < 1 or oh < 1:
            raise RuntimeError('output must be at least 1x1')
        cl = c > 1 and x.stride(1) == 1                          # same choice as x.suggest_memory_format()
        y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last if cl else torch.contiguous_format)
        with torch.cuda.device(x.device):
            st = _lib.load().gt_upfirdn2d(_lib.ptr(x), _lib.ptr(f), _lib.ptr(y), _lib.dtype_code(x), n, c, h, w, *x.stride(),
                                          fh, fw, f.stride(0), f.stride(1), oh, ow, *y.stride(),
                                          int(upx), int(upy), int(downx), int(downy), int(padx0), int(pady0), int(bool(flip)),
                                          float(gain), _lib.stream_of(x))
        _lib.check(st, 'upfirdn2d')
        _lib.count_launch()
        return y


def _is_dense(t):
    if t.is_contiguous():
        return True
    if t.ndim == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return True
    # generic test: sort by stride and check the strides tile the storage without gaps or overlap
    dims = sorted((s, n) for s, n in zip(t.stride(), t.shape) if n > 1)
    expect = 1
    for s, n in dims:
        if s != expect:
            return False
        expect *= n
    return True


_PLUGINS = {'bias_act_plugin': _BiasActPlugin, 'upfirdn2d_plugin': _Upfirdn2dPlugin}


def get_plugin(module_name, sources=None, headers=None, source_dir=None, **build_kwargs):
    """Return the prebuilt plugin object for `module_name`.  `sources`, `headers`, `source_dir` and the build kwargs
    are accepted for call-site compatibility and ignored (nothing is built at run time)."""
    if module_name not in _PLUGINS:
        raise RuntimeError(f'gan_track_b200 has no plugin named "{module_name}" (available: {sorted(_PLUGINS)}); '
                           'the StyleGAN3-only filtered_lrelu plugin is out of scope (SURVEY.md section 2)')
    _lib.load()     # fail loudly here, like a failed JIT build would
    return _PLUGINS[module_name]
