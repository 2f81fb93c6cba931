"""grid_sample (bilinear, zeros padding, align_corners=False) whose backward is itself differentiable w.r.t.
grad_output -- needed by R1 through the ADA pipe.  Interface of OPS/grid_sample_gradfix.py:23-31 (`enabled`,
`grid_sample(input, grid)`).  The reference resolves `aten::grid_sampler_2d_backward` through
`torch._C._jit_get_operation`, which no longer returns a callable on torch 2.x; this module calls
the op through `torch.ops.aten` instead.

Sampling is linear in `input`, so the gradient w.r.t. `input` is the adjoint (scatter) of the sampling operator and the
derivative of that adjoint w.r.t. its cotangent is the sampling operator again: `_Adjoint.backward` re-applies `_Sample`
with the same grid.  Second derivatives w.r.t. the grid are not provided (the pipe's grids carry no gradient).
"""
import torch

enabled = True

_BILINEAR, _ZEROS = 0, 0        # aten enum values of interpolation mode / padding mode


class _Sample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, grid):
        assert image.ndim == 4 and grid.ndim == 4
        ctx.save_for_backward(image, grid)
        return torch.nn.functional.grid_sample(image, grid, mode='bilinear', padding_mode='zeros', align_corners=False)

    @staticmethod
    def backward(ctx, cotangent):
        return _Adjoint.apply(cotangent, *ctx.saved_tensors)


class _Adjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cotangent, image, grid):
        ctx.save_for_backward(grid)
        wanted = list(ctx.needs_input_grad[1:3])
        return tuple(torch.ops.aten.grid_sampler_2d_backward(cotangent, image, grid, _BILINEAR, _ZEROS, False, wanted))

    @staticmethod
    def backward(ctx, d_image_grad, d_grid_grad):
        assert not ctx.needs_input_grad[2]
        grid, = ctx.saved_tensors
        return (_Sample.apply(d_image_grad, grid) if ctx.needs_input_grad[0] else None), None, None


def grid_sample(input, grid):
    return _Sample.apply(input, grid)
