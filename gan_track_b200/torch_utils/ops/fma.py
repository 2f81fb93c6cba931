"""fma(a, b, c) = a * b + c with broadcast-aware gradients (reference: OPS/fma.py:15-58).

One `addcmul` forward; the backward sends each cotangent back to the shape of its operand with `Tensor.sum_to_size`, i.e. it
sums over exactly the dimensions that broadcasting expanded ([N,C,H,W] * [N,C,1,1] + [N,1,H,W] in modulated_conv2d)."""
import torch


class _FusedMultiplyAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        ctx.save_for_backward(a, b)
        ctx.shapes = (a.shape, b.shape, c.shape)
        return torch.addcmul(c, a, b)

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        cotangents = (lambda: dout * b, lambda: dout * a, lambda: dout)
        return tuple(g().sum_to_size(shape) if need else None for g, shape, need in zip(cotangents, ctx.shapes, ctx.needs_input_grad))


def fma(a, b, c):
    return _FusedMultiplyAdd.apply(a, b, c)
