"""conv2d / conv_transpose2d with gradients of arbitrary order, behind the reference's interface
(OPS/conv2d_gradfix.py:22-45: `enabled`, `weight_gradients_disabled`, `no_weight_gradients()`, `conv2d`,
`conv_transpose2d`).

Every derivative of a convolution is again a convolution, so one autograd.Function family closes the algebra:
    y  = conv(x, w)                       forward
    dx = conv^T(dy, w)                    data gradient      -> the same Function with `transpose` flipped
    dw = wgrad(dy, x)                     weight gradient    -> its own Function whose backward is conv / conv^T again
which is what lets R1 (grad-of-grad through D) and the path-length regulariser (grad-of-grad through G) run on the
custom kernels.  The three primitives are provided by `conv_backend` (csrc/conv_igemm.cu: tcgen05 implicit GEMM for
the fp16 channels-last shapes of the StyleGAN2 path; see that module for the coverage table).

`no_weight_gradients()` keeps the reference's meaning: inside it, backward passes skip dw (used by the R1 and
path-length passes, S3/training/loss.py:90, 126).
"""
import contextlib

import torch

from ... import _lib
from . import conv_backend

enabled = True                      # kept for interface compatibility; this package has no alternative path
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients(disable=True):
    global weight_gradients_disabled
    old = weight_gradients_disabled
    if disable:
        weight_gradients_disabled = True
    try:
        yield
    finally:
        weight_gradients_disabled = old


def _pair(v):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    assert len(v) == 2 and all(isinstance(i, int) for i in v)
    return v


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    _lib.require_cuda(input, 'conv2d input')
    assert _pair(dilation) == (1, 1), 'dilation is not used on the StyleGAN2 path and is not supported'
    y = _conv_fn(False, tuple(weight.shape), _pair(stride), _pair(padding), (0, 0), groups).apply(input, weight)
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1)
    return y


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    _lib.require_cuda(input, 'conv_transpose2d input')
    assert _pair(dilation) == (1, 1), 'dilation is not used on the StyleGAN2 path and is not supported'
    y = _conv_fn(True, tuple(weight.shape), _pair(stride), _pair(padding), _pair(output_padding), groups).apply(input, weight)
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1)
    return y


_cache = dict()


def _conv_fn(transpose, weight_shape, stride, padding, output_padding, groups):
    key = (transpose, weight_shape, stride, padding, output_padding, groups)
    if key in _cache:
        return _cache[key]
    kh, kw = weight_shape[2:]
    assert all(s >= 1 for s in stride) and all(p >= 0 for p in padding)
    assert all(0 <= output_padding[i] < max(stride[i], 1) for i in range(2))
    cfg = dict(stride=stride, padding=padding, groups=groups)

    def grad_output_padding(input_shape, output_shape):
        """output_padding of the transposed conv that maps dy back onto the input extent."""
        if transpose:
            return (0, 0)
        return tuple(input_shape[i + 2] - (output_shape[i + 2] - 1) * stride[i] - (1 - 2 * padding[i]) - (weight_shape[i + 2] - 1)
                     for i in range(2))

    class Conv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            assert w.shape == weight_shape
            ctx.save_for_backward(x if w.requires_grad else _empty(x), w)
            ctx.x_shape = x.shape
            return conv_backend.conv_forward(x, w, transpose=transpose, output_padding=output_padding, **cfg)

        @staticmethod
        def backward(ctx, dy):
            x, w = ctx.saved_tensors
            dx = dw = None
            if ctx.needs_input_grad[0]:
                op = grad_output_padding(ctx.x_shape, dy.shape)
                dx = _conv_fn(not transpose, weight_shape, stride, padding, op, groups).apply(dy, w)
                assert dx.shape == ctx.x_shape
            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                dw = Conv2dGradWeight.apply(dy, x)
                assert dw.shape == weight_shape
            return dx, dw

    class Conv2dGradWeight(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x):
            ctx.save_for_backward(dy if x.requires_grad else _empty(dy), x if dy.requires_grad else _empty(x))
            ctx.dy_shape, ctx.x_shape = dy.shape, x.shape
            return conv_backend.conv_wgrad(dy, x, weight_shape, transpose=transpose, output_padding=output_padding, **cfg)

        @staticmethod
        def backward(ctx, d_dw):
            dy, x = ctx.saved_tensors
            d_dy = d_x = None
            if ctx.needs_input_grad[0]:
                d_dy = Conv2d.apply(x, d_dw)
                assert d_dy.shape == ctx.dy_shape
            if ctx.needs_input_grad[1]:
                op = grad_output_padding(ctx.x_shape, ctx.dy_shape)
                d_x = _conv_fn(not transpose, weight_shape, stride, padding, op, groups).apply(dy, d_dw)
                assert d_x.shape == ctx.x_shape
            return d_dy, d_x

    _cache[key] = Conv2d
    return Conv2d


def _empty(like):
    return torch.empty([0], dtype=like.dtype, device=like.device)
