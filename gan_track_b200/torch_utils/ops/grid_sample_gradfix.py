"""grid_sample (bilinear, zeros padding, align_corners=False) whose backward is itself differentiable w.r.t.
grad_output -- needed by R1 through the ADA pipe.  Interface of OPS/grid_sample_gradfix.py:23-31 (`enabled`,
`grid_sample(input, grid)`).  The reference resolves `aten::grid_sampler_2d_backward` through
`torch._C._jit_get_operation`, which no longer returns a callable on torch 2.x; this module calls
the op through `torch.ops.aten` instead.
"""
import torch

enabled = True


def grid_sample(input, grid):
    return _GridSample2dForward.apply(input, grid)


class _GridSample2dForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, grid):
        assert input.ndim == 4 and grid.ndim == 4
        out = torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros', align_corners=False)
        ctx.save_for_backward(input, grid)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, grid = ctx.saved_tensors
        grad_input, grad_grid = _GridSample2dBackward.apply(grad_output, input, grid)
        return grad_input, grad_grid


class _GridSample2dBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, input, grid):
        mask = [ctx.needs_input_grad[1], ctx.needs_input_grad[2]]
        grad_input, grad_grid = torch.ops.aten.grid_sampler_2d_backward(grad_output, input, grid, 0, 0, False, mask)
        ctx.save_for_backward(grid)
        return grad_input, grad_grid

    @staticmethod
    def backward(ctx, grad2_grad_input, grad2_grad_grid):
        grid, = ctx.saved_tensors
        grad2_grad_output = None
        if ctx.needs_input_grad[0]:
            grad2_grad_output = _GridSample2dForward.apply(grad2_grad_input, grid)
        assert not ctx.needs_input_grad[2]
        return grad2_grad_output, None, None
