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

    Conv2d.GradWeight = Conv2dGradWeight
    Conv2d.grad_output_padding = staticmethod(grad_output_padding)
    _cache[key] = Conv2d
    return Conv2d


def _empty(like):
    return torch.empty([0], dtype=like.dtype, device=like.device)


# ---- convolution with the layer's bias_act fused into the kernel epilogue ----------------------------------------------

_ACT_CODE = {'linear': 1, 'lrelu': 3}
import os as _os
# True: fuse wherever the tcgen05 kernels take the convolution; False: never; 'auto' (default): where it was measured to win on B200
# (tools/bench_fused_epilogue.py -> profiles/r02_fused_epilogue.txt; us per layer forward at batch 32, separate -> fused):
#   64 -> 64 3x3 @256^2 (row-streaming kernel)  205 -> 139      128 -> 128 3x3 @128^2  159 -> 133      256 -> 256 3x3 @64^2  133 -> 116
#   512 -> 512 3x3 @32^2  132 -> 123            1x1 skip 64 -> 128 @128^2  112 -> 83          64 -> 128 stride 2 @257  182 -> 174
# and NOT on the stride-2 layers with 128 / 256 input channels (122 -> 141, 103 -> 111: the per-tap kernel's epilogue is not overlapped
# with the next tile's MMAs).  The fused mode of the row-streaming and CTA-pair kernels runs EIGHT epilogue warps: with four, the extra
# per-element work made the narrow layers epilogue-paced (64 channels: 282 us fused).  GT_FUSE_BIAS_ACT=0 / 1 / auto (or the attribute).
_env = _os.environ.get('GT_FUSE_BIAS_ACT', 'auto')
fuse_bias_act = True if _env == '1' else (False if _env == '0' else 'auto')


fuse_residual = True      # False: a residual (`addend`) is added by its own pass after the (possibly fused) bias_act -- A/B switch


def _fuse_profitable(input, weight, stride):
    cout, cin, kh, kw = weight.shape
    if stride == (1, 1):
        return True
    return cin <= 64 or cin >= 512


def conv2d_bias_act(input, weight, bias, act='linear', alpha=None, gain=None, clamp=None, stride=1, padding=0, addend=None):
    """clamp(act(conv2d(input, weight) + bias) * gain), the tail of Conv2dLayer.forward (S3/training/networks_stylegan2.py:173-177).
    When the tcgen05 kernel takes the convolution (fp16 channels-last, channels multiples of 64) the bias / activation /
    gain / clamp run in its epilogue -- the activation tensor is written once instead of written, re-read and re-written --
    otherwise this is exactly `bias_act(conv2d(...))`.  `addend` (optional, the shape of the result): added to the result, in the
    same epilogue when fused -- DiscriminatorBlock's `y.add_(x)` (:636) folded into the skip convolution.  Gradients of any order: backward = bias_act's gradient Function
    followed by the convolution's, both differentiable."""
    from . import bias_act as bias_act_mod
    from . import conv_igemm
    stride, padding = _pair(stride), _pair(padding)
    fusable = (fuse_bias_act is not False and conv_backend.allow_igemm and act in _ACT_CODE and input.is_cuda and input.dtype == torch.float16
               and (fuse_bias_act is True or _fuse_profitable(input, weight, stride))
               and conv_igemm.covered(input, weight, False, (0, 0), stride, padding, 1) and (bias is None or bias.numel() % 8 == 0))
    if addend is not None and not fuse_residual:
        return conv2d_bias_act(input, weight, bias, act=act, alpha=alpha, gain=gain, clamp=clamp, stride=stride, padding=padding).add_(addend)
    if addend is not None and fusable:
        fusable = 'y' not in bias_act_mod.activation_funcs[act].ref
    if addend is not None and fusable:
        OH, OW = conv_igemm.out_size(input.shape[2], input.shape[3], weight.shape[2], weight.shape[3], stride[0], padding[0], False)
        dense_cl = addend.dim() == 4 and addend.stride() == (OH * OW * weight.shape[0], 1, OW * weight.shape[0], weight.shape[0])
        fusable = (addend.dtype == torch.float16 and tuple(addend.shape) == (input.shape[0], weight.shape[0], OH, OW) and dense_cl
                   and addend.data_ptr() % 16 == 0)
    if not fusable:
        y = conv2d(input, weight, stride=stride, padding=padding)
        y = bias_act_mod.bias_act(y, bias, act=act, alpha=alpha, gain=gain, clamp=clamp)
        return y if addend is None else y.add_(addend)
    bias_act_mod._init()
    BA = bias_act_mod._bias_act_cuda(dim=1, act=act, alpha=alpha, gain=gain, clamp=clamp)
    spec, alpha_f, gain_f, clamp_f, trivial = BA.cfg
    Conv = _conv_fn(False, tuple(weight.shape), stride, padding, (0, 0), 1)
    return _fused_fn(Conv, BA, stride, padding).apply(input, weight, bias, addend)


_fused_cache = dict()


def _fused_fn(Conv, BA, stride, padding):
    key = (Conv, BA)
    if key in _fused_cache:
        return _fused_cache[key]
    from . import bias_act as bias_act_mod
    from . import conv_igemm
    spec, alpha, gain, clamp, trivial = BA.cfg

    class ConvBiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b, addend):
            bb = b.to(torch.float16).contiguous() if b is not None else None
            y = conv_igemm.igemm_forward(x, w, transpose=False, output_padding=(0, 0), stride=stride, padding=padding, groups=1,
                                         epilogue=(bb, spec.cuda_idx, alpha, gain, clamp, addend))
            assert y is not None
            conv_backend.stats['igemm'] += 1
            # like the reference's bias_act, `linear` saves no output (callers may then update it in place, e.g. y.add_(x) in
            # DiscriminatorBlock.forward, S3/training/networks_stylegan2.py:636).  With a residual the stored tensor is act(...) + addend,
            # which is not the activation's output: only `linear` (nothing saved) may carry one.
            assert addend is None or 'y' not in spec.ref, 'a fused residual needs an activation whose gradient does not read its output'
            ctx.save_for_backward(x if w.requires_grad else _empty(x), w, y if 'y' in spec.ref else _empty(y))
            ctx.x_shape = x.shape
            ctx.has_bias = b is not None
            return y

        @staticmethod
        def backward(ctx, dy):
            x, w, y = ctx.saved_tensors
            dx = dw = db = None
            want_db = ctx.has_bias and ctx.needs_input_grad[2]
            dy = dy.contiguous(memory_format=torch.channels_last)
            # gradient of bias_act from the saved OUTPUT (lrelu / linear need nothing else, OPS/bias_act.py:151-154)
            # (`linear` saves no output in the reference, so its clamp does not mask the gradient: pass y only where the reference does)
            ysave = y if y.numel() else None
            if want_db and not torch.is_grad_enabled():
                dpre, db = bias_act_mod._fused_bwd(dy, ysave, 1, spec, alpha, gain, clamp)      # dx and db in one pass
            else:
                e = bias_act_mod._empty
                dpre = dy if trivial else BA.Grad.apply(dy, e, e, ysave if ysave is not None else e)
                if want_db:
                    db = dpre.sum([0, 2, 3])
            if ctx.needs_input_grad[0]:
                op = Conv.grad_output_padding(ctx.x_shape, dpre.shape)
                dx = _conv_fn(True, tuple(w.shape), stride, padding, op, 1).apply(dpre, w)
            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                dw = Conv.GradWeight.apply(dpre, x)
            return dx, dw, db, (dy if ctx.needs_input_grad[3] else None)      # the residual passes the gradient through

    _fused_cache[key] = ConvBiasAct
    return ConvBiasAct
