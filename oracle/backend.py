"""ORACLE harness (test infrastructure): run the product's HOST code (conv2d_resample case split, modulated_conv2d,
networks, loss, AugmentPipe, training step) on the CPU with the oracle's primitive ops in place of the CUDA kernels.

The product package has no CPU path; its primitive ops raise on CPU tensors.  Inside `oracle_ops()` the five primitive
entry points are swapped for the restatements in `oracle/ops_ref.py`:
    bias_act.bias_act, upfirdn2d.upfirdn2d, conv2d_gradfix.conv2d, conv2d_gradfix.conv_transpose2d, fma.fma
(grid_sample_gradfix is built from device-agnostic aten ops and is left alone).  Nothing in `gan_track_b200/` imports
this module; only tests, smoke() and bench.py's CPU legs do.
"""
import contextlib

from . import ops_ref


@contextlib.contextmanager
def oracle_ops():
    from gan_track_b200.torch_utils.ops import bias_act, conv2d_gradfix, fma, upfirdn2d
    saved = [
        (bias_act, 'bias_act', bias_act.bias_act),
        (upfirdn2d, 'upfirdn2d', upfirdn2d.upfirdn2d),
        (conv2d_gradfix, 'conv2d', conv2d_gradfix.conv2d),
        (conv2d_gradfix, 'conv_transpose2d', conv2d_gradfix.conv_transpose2d),
        (fma, 'fma', fma.fma),
    ]
    bias_act.bias_act = ops_ref.bias_act
    upfirdn2d.upfirdn2d = ops_ref.upfirdn2d
    conv2d_gradfix.conv2d = ops_ref.conv2d
    conv2d_gradfix.conv_transpose2d = ops_ref.conv_transpose2d
    fma.fma = ops_ref.fma
    try:
        yield
    finally:
        for mod, name, fn in saved:
            setattr(mod, name, fn)
