"""Fully-connected layers at training batch sizes (csrc/fc.cu): `linear(x, weight, bias, weight_gain, bias_gain)` =
`addmm(bias * bias_gain, x, (weight * weight_gain).t())`, the body of FullyConnectedLayer.forward
(S3/training/networks_stylegan2.py:115-126), with gradients of arbitrary order.

Every derivative of the three primitives is another of the three,
    fwd(x, w, b)   = wg x w^T + bg b          dgrad(dy, w) = wg dy w          wgrad(dy, x) = (wg dy^T x, bg sum_m dy)
so one autograd.Function family closes the algebra -- the path-length regulariser differentiates the style affines'
backward and R1 the discriminator epilogue's (S3/training/loss.py:85-100, 120-133).

`applicable(x, weight)`: CUDA, fp32, contiguous 2-D, batch <= 64, in_features a multiple of 4; callers keep the reference's op
sequence otherwise.
"""
import torch

from ... import _lib


def applicable(x, weight):
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.ndim == 2 and x.dtype == torch.float32 and weight.dtype == torch.float32
            and 1 <= x.shape[0] <= 64 and x.shape[1] % 4 == 0 and x.shape[1] == weight.shape[1])


def _c(t):
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _fwd(x, w, b, wg, bg):
    x, w = _c(x), _c(w)
    M, I = x.shape
    O = w.shape[0]
    y = torch.empty([M, O], dtype=torch.float32, device=x.device)
    bb = b.contiguous() if b is not None else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().gt_fc_fwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bb), _lib.ptr(y), M, I, O, wg, bg, _lib.stream_of(x)), 'gt_fc_fwd')
    _lib.count_launch()
    return y


def _dgrad(dy, w, wg):
    dy, w = dy.contiguous(), w.contiguous()
    M, O = dy.shape
    I = w.shape[1]
    dx = torch.empty([M, I], dtype=torch.float32, device=dy.device)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().gt_fc_dgrad(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(dx), M, I, O, wg, _lib.stream_of(dy)), 'gt_fc_dgrad')
    _lib.count_launch()
    return dx


def _wgrad(dy, x, wg, bg, want_db):
    dy, x = dy.contiguous(), x.contiguous()
    M, O = dy.shape
    I = x.shape[1]
    dw = torch.empty([O, I], dtype=torch.float32, device=dy.device)
    db = torch.empty([O], dtype=torch.float32, device=dy.device) if want_db else None
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().gt_fc_wgrad(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), _lib.ptr(db), M, I, O, wg, bg, _lib.stream_of(dy)), 'gt_fc_wgrad')
    _lib.count_launch()
    return dw, db


_cache = {}


def _family(wg, bg, has_bias):
    key = (wg, bg, has_bias)
    if key in _cache:
        return _cache[key]

    class Fwd(torch.autograd.Function):            # y = wg x w^T (+ bg b)
        @staticmethod
        def forward(ctx, x, w, b):
            ctx.save_for_backward(x, w)
            return _fwd(x, w, b, wg, bg)

        @staticmethod
        def backward(ctx, dy):
            x, w = ctx.saved_tensors
            dx = dw = db = None
            if ctx.needs_input_grad[0]:
                dx = Dgrad.apply(dy, w)
            if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
                dw, db = Wgrad.apply(dy, x)
                if not ctx.needs_input_grad[1]:
                    dw = None
                if not (has_bias and ctx.needs_input_grad[2]):
                    db = None
            return dx, dw, db

    class Dgrad(torch.autograd.Function):          # dx = wg dy w
        @staticmethod
        def forward(ctx, dy, w):
            ctx.save_for_backward(dy, w)
            return _dgrad(dy, w, wg)

        @staticmethod
        def backward(ctx, ggx):
            dy, w = ctx.saved_tensors
            d_dy = d_w = None
            if ctx.needs_input_grad[0]:
                d_dy = _family(wg, bg, False)[0].apply(ggx, w, None)          # wg ggx w^T
            if ctx.needs_input_grad[1]:
                d_w, _ = _family(wg, bg, False)[2].apply(dy, ggx)             # wg dy^T ggx
            return d_dy, d_w

    class Wgrad(torch.autograd.Function):          # dw = wg dy^T x, db = bg sum_m dy
        @staticmethod
        def forward(ctx, dy, x):
            ctx.save_for_backward(dy, x)
            dw, db = _wgrad(dy, x, wg, bg, has_bias)
            if db is None:
                db = torch.zeros([0], dtype=torch.float32, device=dy.device)
                ctx.mark_non_differentiable(db)
            return dw, db

        @staticmethod
        def backward(ctx, ggw, ggb):
            dy, x = ctx.saved_tensors
            d_dy = d_x = None
            if ctx.needs_input_grad[0]:
                gb = ggb if (has_bias and ggb is not None and ggb.numel()) else None
                if ggw is not None:
                    d_dy = _family(wg, bg, gb is not None)[0].apply(x, ggw, gb)      # wg x ggw^T + bg ggb
                elif gb is not None:
                    d_dy = (gb * bg).unsqueeze(0).expand(dy.shape[0], -1)
            if ctx.needs_input_grad[1] and ggw is not None:
                d_x = Dgrad.apply(dy, ggw)                                           # wg dy ggw
            return d_dy, d_x

    _cache[key] = (Fwd, Dgrad, Wgrad)
    return _cache[key]


def linear(x, weight, bias=None, weight_gain=1.0, bias_gain=1.0):
    """x: [M, I] fp32 (M <= 64), weight: [O, I], bias: [O] or None.  Raises (through the C ABI) for shapes it does not cover;
    check `applicable` first."""
    _lib.require_cuda(x, 'fc input')
    Fwd = _family(float(weight_gain), float(bias_gain), bias is not None)[0]
    return Fwd.apply(x, weight, bias)
