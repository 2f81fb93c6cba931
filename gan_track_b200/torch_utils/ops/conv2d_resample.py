"""conv2d_resample: 2-D convolution with optional FIR up/down-sampling, reference signature
(OPS/conv2d_resample.py:46).  Padding is relative to the upsampled image and applied once.

Which kernels run for each case on the StyleGAN2 path:
    up=1 down=1            -> conv (3x3 pad 1 or 1x1)
    up=2 (3x3)             -> transposed stride-2 conv, then 4x4 FIR with gain 4
    down=2 (3x3)           -> 4x4 FIR (pad 2), then stride-2 conv
    down=2 (1x1)           -> 4x4 FIR with decimation, then 1x1 conv
    up=2 (1x1)             -> 1x1 conv, then 4x4 FIR upsampling
"""
import torch

from .. import misc
from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _get_filter_size, _parse_padding


def _get_weight_shape(w):
    return [int(sz) for sz in w.shape]


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True, epilogue=None):
    """conv2d / conv_transpose2d; `flip_weight=False` means true convolution (taps reversed).  `epilogue`: dict(b, act, gain,
    clamp) -- the bias_act that follows the convolution, fused into the kernel where it can be (conv2d_gradfix.conv2d_bias_act)."""
    _oc, _icg, kh, kw = _get_weight_shape(w)
    if not flip_weight and (kw > 1 or kh > 1):
        w = w.flip([2, 3])
    if epilogue is not None:
        assert not transpose and groups == 1
        return conv2d_gradfix.conv2d_bias_act(x, w, epilogue['b'], act=epilogue['act'], gain=epilogue['gain'], clamp=epilogue['clamp'],
                                              stride=stride, padding=padding, addend=epilogue.get('addend'))
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


@misc.profiled_function
def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False, epilogue=None):
    """`epilogue` (not in the reference): dict(b, act, gain, clamp[, addend]) applied as bias_act (+ residual) to the result; only accepted when the
    convolution is the LAST kernel of the case split (up == 1, groups == 1), where it is fused into the convolution."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1
    assert isinstance(down, int) and down >= 1
    assert isinstance(groups, int) and groups >= 1
    oc, icg, kh, kw = _get_weight_shape(w)
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)

    # The FIR widens the footprint; keep the output aligned with the input grid.
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2

    assert epilogue is None or (up == 1 and groups == 1), 'a fused epilogue needs the convolution to be the last kernel'
    pointwise = (kw == 1 and kh == 1)
    if pointwise and down > 1 and up == 1:          # decimate first: the 1x1 conv then sees 1/4 of the pixels
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight, epilogue=epilogue)

    if pointwise and up > 1 and down == 1:          # convolve first: the 1x1 conv sees the small image
        x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d.upfirdn2d(x=x, f=f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)

    if down > 1 and up == 1:                        # blur, then strided conv
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, stride=down, groups=groups, flip_weight=flip_weight, epilogue=epilogue)

    if up > 1:                                      # transposed strided conv, then blur (and optional decimation)
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, oc // groups, icg, kh, kw).transpose(1, 2).reshape(groups * icg, oc // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv2d_wrapper(x=x, w=w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
        return x

    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return _conv2d_wrapper(x=x, w=w, padding=[py0, px0], groups=groups, flip_weight=flip_weight, epilogue=epilogue)

    # Anything else: explicit pad/upsample, conv, decimate.
    x = upfirdn2d.upfirdn2d(x=x, f=(f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
    if epilogue is not None:
        from . import bias_act
        x = bias_act.bias_act(x, epilogue['b'], act=epilogue['act'], gain=epilogue['gain'], clamp=epilogue['clamp'])
        if epilogue.get('addend') is not None:
            x = x.add_(epilogue['addend'])
    return x
