"""The three convolution primitives behind `conv2d_gradfix`: forward / data-gradient (same kernel family) and
weight-gradient.

Coverage (see DESIGN.md "conv" for the table that is kept current):
  * fp16, channels-last, groups == 1, Cin and Cout multiples of 64: tcgen05 implicit GEMM (csrc/conv_igemm.cu)
    - 3x3 stride 1 pad 1, 1x1 stride 1, 3x3 stride 2 pad 0 (D down path), 3x3 transposed stride 2 (G up path).
  * everything else on the path (fp32 low-resolution layers, 1-channel ToRGB / FromRGB, grouped eval-mode convs,
    weight gradients until the tcgen05 wgrad lands) is routed to the library convolution of the host framework
    (`aten::convolution` -> cuDNN), exactly what the reference itself calls (OPS/conv2d_gradfix.py:40,45).  The
    route taken is counted in `stats` so tests and bench.py can report how many launches were ours.
"""
import torch

from ... import _lib

stats = {'igemm': 0, 'library': 0, 'library_wgrad': 0, 'igemm_wgrad': 0}

# Set by conv_igemm (if the kernels are available) -- callables returning None when a shape is not covered.
_igemm_forward = None
_igemm_wgrad = None

allow_igemm = True
# False: a convolution outside the coverage of this package's kernels raises instead of going to the library (bench.py --no-library and
# the GPU test that pins the 256x256 model to the repository's own kernels); the default keeps the reference's own route (F.conv2d) as fallback
allow_library = True

# fp32 layers must be true fp32 (north star: 1e-5 relative; the reference turns TF32 off in its training loop,
# S3/training/training_loop_mi_multimodal.py:169-170).  The library route would otherwise silently use TF32.
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _aten_conv(x, w, stride, padding, transpose, output_padding, groups):
    return torch.ops.aten.convolution(x, w, None, list(stride), list(padding), [1, 1], transpose, list(output_padding), groups)


def conv_forward(x, w, *, transpose, output_padding, stride, padding, groups):
    _lib.require_cuda(x, 'conv input')
    if allow_igemm and _igemm_forward is not None:
        y = _igemm_forward(x, w, transpose=transpose, output_padding=output_padding, stride=stride, padding=padding, groups=groups)
        if y is not None:
            stats['igemm'] += 1
            return y
    if not allow_library:
        raise RuntimeError(f'conv_backend: no kernel of this package covers conv x{tuple(x.shape)} {x.dtype} w{tuple(w.shape)} stride {tuple(stride)} '
                           f'padding {tuple(padding)} transpose {transpose} groups {groups} and allow_library is False')
    stats['library'] += 1
    return _aten_conv(x, w, stride, padding, transpose, output_padding, groups)


def conv_wgrad(dy, x, weight_shape, *, transpose, output_padding, stride, padding, groups):
    _lib.require_cuda(x, 'conv input')
    if allow_igemm and _igemm_wgrad is not None:
        dw = _igemm_wgrad(dy, x, weight_shape, transpose=transpose, output_padding=output_padding, stride=stride, padding=padding, groups=groups)
        if dw is not None:
            stats['igemm_wgrad'] += 1
            return dw
    if not allow_library:
        raise RuntimeError(f'conv_backend: no weight-gradient kernel of this package covers dy{tuple(dy.shape)} x{tuple(x.shape)} {x.dtype} w{tuple(weight_shape)} '
                           f'stride {tuple(stride)} transpose {transpose} groups {groups} and allow_library is False')
    stats['library_wgrad'] += 1
    w_dummy = torch.empty(weight_shape, dtype=x.dtype, device=x.device)
    _, dw, _ = torch.ops.aten.convolution_backward(dy, x, w_dummy, None, list(stride), list(padding), [1, 1], transpose,
                                                   list(output_padding), groups, [False, True, False])
    return dw


from . import conv_igemm  # noqa: E402  (registers the tcgen05 kernels as the first route)

conv_igemm.install()
