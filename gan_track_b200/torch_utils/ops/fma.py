"""fma(a, b, c) = a * b + c with broadcast-aware gradients (reference: OPS/fma.py:15-58)."""
import torch


def fma(a, b, c):
    return _FusedMultiplyAdd.apply(a, b, c)


class _FusedMultiplyAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        out = torch.addcmul(c, a, b)
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        da = db = dc = None
        if ctx.needs_input_grad[0]:
            da = _unbroadcast(dout * b, a.shape)
        if ctx.needs_input_grad[1]:
            db = _unbroadcast(dout * a, b.shape)
        if ctx.needs_input_grad[2]:
            dc = _unbroadcast(dout, ctx.c_shape)
        return da, db, dc


def _unbroadcast(x, shape):
    """Sum x over the dims that were broadcast to reach x.shape from `shape`."""
    extra = x.ndim - len(shape)
    assert extra >= 0
    dims = [i for i in range(x.ndim) if x.shape[i] > 1 and (i < extra or shape[i - extra] == 1)]
    if dims:
        x = x.sum(dim=dims, keepdim=True)
    if extra:
        x = x.reshape(-1, *x.shape[extra + 1:])
    assert x.shape == shape
    return x
