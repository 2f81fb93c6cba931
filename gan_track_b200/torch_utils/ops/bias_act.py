"""bias_act: y = clamp(act(x + b) * gain), with first and second order gradients.

Same call signature and activation table as the reference (OPS/bias_act.py:21-31, :52-86).  The work is done by
`gt_bias_act` / `gt_bias_act_bwd` (csrc/bias_act.cu) through the plugin object of `custom_ops.get_plugin`.
Differences in execution, not in results:
  * the first-order backward of linear / lrelu produces dx AND db in one pass when no graph is being recorded
    (the reference re-reads dx with `dx.sum(...)`, OPS/bias_act.py:169-170);
  * there is no `ref` implementation here -- CPU tensors raise.  The CPU restatement lives in oracle/ops_ref.py.
"""
import numpy as np
import torch

from ... import _lib
from ...dnnlib import EasyDict
from .. import custom_ops

activation_funcs = {
    'linear':   EasyDict(func=lambda x, **_: x,                                            def_alpha=0,   def_gain=1,          cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     EasyDict(func=lambda x, **_: torch.nn.functional.relu(x),                  def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    EasyDict(func=lambda x, alpha, **_: torch.nn.functional.leaky_relu(x, alpha), def_alpha=0.2, def_gain=np.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     EasyDict(func=lambda x, **_: torch.tanh(x),                                def_alpha=0,   def_gain=1,          cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  EasyDict(func=lambda x, **_: torch.sigmoid(x),                             def_alpha=0,   def_gain=1,          cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      EasyDict(func=lambda x, **_: torch.nn.functional.elu(x),                   def_alpha=0,   def_gain=1,          cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     EasyDict(func=lambda x, **_: torch.nn.functional.selu(x),                  def_alpha=0,   def_gain=1,          cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': EasyDict(func=lambda x, **_: torch.nn.functional.softplus(x),              def_alpha=0,   def_gain=1,          cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    EasyDict(func=lambda x, **_: torch.sigmoid(x) * x,                         def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}

_plugin = None
_empty = torch.empty([0])


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin(module_name='bias_act_plugin', sources=['bias_act.cu'], headers=['gt_common.cuh'])
    return True


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """Fused bias + activation + gain + clamp.  Arguments as in the reference (OPS/bias_act.py:52-81).
    `impl` is accepted for signature compatibility; only 'cuda' exists in this package."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl != 'cuda':
        raise NotImplementedError("gan_track_b200 ships only the CUDA implementation; the 'ref' restatement is oracle/ops_ref.py")
    _lib.require_cuda(x, 'bias_act input')
    _init()
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)


_cache = dict()


def _mem_format(t):
    return torch.channels_last if t.ndim == 4 and t.stride(1) == 1 and t.shape[1] > 1 else torch.contiguous_format


def _fused_bwd(dy, y, dim, spec, alpha, gain, clamp):
    """dx and db from one pass (csrc/bias_act.cu: gt_bias_act_bwd).  dy, y dense with identical layout."""
    C = dy.shape[dim]
    stride = dy.stride(dim)
    if stride == 1:                       # channels-last 4-D or [rows, C] with dim last
        outer, inner = dy.numel() // C, 1
    else:                                 # plain contiguous: [outer, C, inner]
        inner = stride
        outer = dy.numel() // (C * inner)
    lib = _lib.load()
    ws_n = int(lib.gt_bias_act_bwd_workspace(outer, C, inner))
    ws = torch.empty([ws_n], dtype=torch.float32, device=dy.device)
    dx = torch.empty_like(dy)
    db = torch.empty([C], dtype=torch.float32, device=dy.device)
    with torch.cuda.device(dy.device):
        st = lib.gt_bias_act_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(dx), _lib.ptr(db), _lib.ptr(ws), ws_n, _lib.dtype_code(dy),
                                 spec.cuda_idx, alpha, gain, clamp, outer, C, inner, _lib.stream_of(dy))
    _lib.check(st, 'bias_act_bwd')
    _lib.count_launch(2)
    return dx, db.to(dy.dtype)


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    key = (dim, act, alpha, gain, clamp)
    if key in _cache:
        return _cache[key]

    trivial = (act == 'linear' and gain == 1 and clamp < 0)
    # What is saved follows the reference exactly (OPS/bias_act.py:151-154): `linear` saves nothing, so its clamp does
    # not mask the gradient (the plugin sees yref = 0, OPS/bias_act.cu:47,137-142); lrelu saves y.
    save_x = 'x' in spec.ref or spec.has_2nd_grad
    save_y = 'y' in spec.ref

    class BiasActCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            ctx.memory_format = _mem_format(x)
            x = x.contiguous(memory_format=ctx.memory_format)
            b = b.contiguous() if b is not None else _empty
            y = x
            if not trivial or b is not _empty:
                y = _plugin.bias_act(x, b, _empty, _empty, _empty, 0, dim, spec.cuda_idx, alpha, gain, clamp)
            ctx.save_for_backward(x if save_x else _empty, b if save_x else _empty, y if save_y else _empty)
            ctx.has_bias = b is not _empty
            return y

        @staticmethod
        def backward(ctx, dy):
            dy = dy.contiguous(memory_format=ctx.memory_format)
            x, b, y = ctx.saved_tensors
            dx = db = None
            want_dx, want_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            if not (want_dx or want_db):
                return None, None
            fusable = (want_db and act in ('linear', 'lrelu') and not trivial
                       and dy.dtype in (torch.float16, torch.float32) and dy.numel() > 0
                       and (dy.stride(dim) == 1 or dy.is_contiguous()))
            if fusable and not torch.is_grad_enabled():
                dx, db = _fused_bwd(dy, y if y.numel() else None, dim, spec, alpha, gain, clamp)
                return dx, db
            if fusable:
                # under create_graph (R1 differentiates the discriminator's backward, S3/training/loss.py:120-133) the same fused pass as a
                # differentiable Function: the bias gradient is never used by that pass, and `dx.sum(...)` over the activation tensor
                # (what the reference runs, OPS/bias_act.py:169-170) would cost as much as the gradient itself
                return BiasActCudaGradFused.apply(dy, y)
            dx = dy
            if not trivial:
                dx = BiasActCudaGrad.apply(dy, x, b, y)
            if want_db:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    class BiasActCudaGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            ctx.memory_format = _mem_format(dy)
            dx = _plugin.bias_act(dy, b, x, y, _empty, 1, dim, spec.cuda_idx, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else _empty, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            d_dx = d_dx.contiguous(memory_format=ctx.memory_format)
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActCudaGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _plugin.bias_act(d_dx, b, x, y, dy, 2, dim, spec.cuda_idx, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    class BiasActCudaGradFused(torch.autograd.Function):
        """(dy, y) -> (dx, db) in one pass for linear / lrelu (no second derivative of the activation): dx = dy * slope(y), db = sum dx.  Its
        backward is the gradient Function again, applied to d_dx + d_db (broadcast)."""

        @staticmethod
        def forward(ctx, dy, y):
            ctx.memory_format = _mem_format(dy)
            dx, db = _fused_bwd(dy, y if y.numel() else None, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(y)
            ctx.dx_shape, ctx.dx_dtype = tuple(dy.shape), dy.dtype
            ctx.set_materialize_grads(False)
            return dx, db

        @staticmethod
        def backward(ctx, d_dx, d_db):
            y, = ctx.saved_tensors
            if not ctx.needs_input_grad[0] or (d_dx is None and d_db is None):
                return None, None
            t = d_dx.contiguous(memory_format=ctx.memory_format) if d_dx is not None else None
            if d_db is not None:
                bb = d_db.to(ctx.dx_dtype).reshape([-1 if i == dim else 1 for i in range(len(ctx.dx_shape))])
                t = (t + bb) if t is not None else bb.expand(ctx.dx_shape).contiguous(memory_format=ctx.memory_format)
            return BiasActCudaGrad.apply(t, _empty, _empty, y), None

    BiasActCuda.Grad = BiasActCudaGrad          # used by the fused conv + bias_act op (conv2d_gradfix.conv2d_bias_act)
    BiasActCuda.cfg = (spec, alpha, gain, clamp, trivial)
    _cache[key] = BiasActCuda
    return BiasActCuda
