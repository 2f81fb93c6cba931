"""Binding of the tcgen05 implicit-GEMM convolution (csrc/conv_igemm.cu, csrc/conv_wgrad.cu) behind `conv_backend`.

`igemm_forward` / `igemm_wgrad` return None when a call is outside the kernels' coverage, in which case
`conv_backend` takes the library route the reference itself uses.  Coverage: fp16, groups = 1, Cin and Cout multiples
of 64, kernels of at most 9 taps, stride 1 (any pad), stride 2, transposed stride 1, transposed stride 2 with pad 0 and
no output padding.  Tensors are consumed channels-last (NHWC); an NCHW input is re-laid-out first.
"""
import torch

from ... import _lib

enabled = True

# {(kind, N, Cin, Cout, H, W, k, stride, transpose): calls} when a dict is installed (bench.py: which convolution shape carries the
# step); Trainer snapshots it around a graph capture so that replays are accounted for.
call_log = None


def _log(kind, N, cin, cout, H, W, k, stride, transpose):
    if call_log is not None:
        key = (kind, int(N), int(cin), int(cout), int(H), int(W), int(k), int(stride), bool(transpose))
        call_log[key] = call_log.get(key, 0) + 1


def _nhwc(t):
    """t as an NHWC-strided tensor whose strides the TMA unit accepts."""
    n, c, h, w = t.shape
    s = t.stride()
    ok = s[1] == 1 and s[3] == c and s[2] == w * c and s[0] == h * w * c and t.data_ptr() % 16 == 0
    if not ok:
        t = t.contiguous(memory_format=torch.channels_last)
        if t.stride(1) != 1 or t.stride(3) != c:            # size-1 dims can leave ambiguous strides; force them
            t = t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    return t


def covered(x, w, transpose, output_padding, stride, padding, groups):
    if not enabled or x.dtype != torch.float16 or w.dtype != torch.float16 or not x.is_cuda or groups != 1:
        return False
    if stride[0] != stride[1] or padding[0] != padding[1] or tuple(output_padding) != (0, 0) or stride[0] not in (1, 2):
        return False
    kh, kw = w.shape[2:]
    cin, cout = (w.shape[0], w.shape[1]) if transpose else (w.shape[1], w.shape[0])
    if x.shape[1] != cin or cin % 64 != 0 or cout % 64 != 0 or kh * kw > 9 or padding[0] >= 8:
        return False
    if transpose and stride[0] == 2 and padding[0] != 0:
        return False
    if min(x.shape) == 0:
        return False
    return True


def out_size(H, W, kh, kw, stride, pad, transpose):
    if transpose:
        return (H - 1) * stride - 2 * pad + kh, (W - 1) * stride - 2 * pad + kw
    return (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1


def pack_weight(w, transpose):
    """[KH*KW][Cout][Cin] fp16, the layout the kernel's weight TMA reads."""
    lib = _lib.load()
    kh, kw = w.shape[2:]
    cin, cout = (w.shape[0], w.shape[1]) if transpose else (w.shape[1], w.shape[0])
    s = w.stride()
    s_co, s_ci = (s[1], s[0]) if transpose else (s[0], s[1])
    packed = torch.empty([kh * kw, cout, cin], dtype=torch.float16, device=w.device)
    _lib.check(lib.gt_conv_pack_weight_f16(_lib.ptr(w), s_co, s_ci, s[2], s[3], cout, cin, kh, kw, _lib.ptr(packed), _lib.stream_of(w)),
               'gt_conv_pack_weight_f16')
    _lib.count_launch()
    return packed


# fp32 layers on the tensor cores as fp16 x 3 (csrc/conv_f16x3.cu): each fp32 operand is scaled by a power of two, split into
# hi + lo fp16 halves and the three significant products run as ONE fp16 implicit GEMM over a 3x wider channel axis with fp32
# output, K-split by kernel row so that no stored value sees more than ~300 truncating accumulator updates; partial sums and the
# rescale are applied in fp32.  Accuracy ~1e-6 relative (tests/test_gpu_conv_igemm.py), i.e. tighter than the library's own
# fp32 kernels at 512 channels; True by default -- the fp32 blocks no longer call the library.
f16x3_enabled = True


def _pad64(c):
    return (c + 63) // 64 * 64


def covered_fp32(x, w, transpose, output_padding, stride, padding, groups):
    if not (enabled and f16x3_enabled) or x.dtype != torch.float32 or w.dtype != torch.float32 or not x.is_cuda or groups != 1:
        return False
    if stride[0] != stride[1] or padding[0] != padding[1] or tuple(output_padding) != (0, 0) or stride[0] not in (1, 2):
        return False
    kh, kw = w.shape[2:]
    cin, cout = (w.shape[0], w.shape[1]) if transpose else (w.shape[1], w.shape[0])
    if x.shape[1] != cin or kh * kw > 9 or kh > 4 or padding[0] >= 8 or cin < 16 or cout < 16:
        return False            # 1-channel ToRGB / FromRGB have their own kernels (ops/rgb.py); tiny channel counts are not worth the padding
    if transpose and stride[0] == 2 and padding[0] != 0:
        return False
    return min(x.shape) > 0


def _amax(t):
    """max|t| as an fp32 bit pattern in a device uint32 (one launch + a 4-byte memset; order independent)."""
    lib = _lib.load()
    out = torch.empty([1], dtype=torch.int32, device=t.device)
    sh = list(t.shape) + [1] * (4 - t.ndim)
    st = list(t.stride()) + [0] * (4 - t.ndim)
    _lib.check(lib.gt_f16x3_amax(_lib.ptr(t), st[0], st[1], st[2], st[3], sh[0], sh[1], sh[2], sh[3], _lib.ptr(out), _lib.stream_of(t)), 'gt_f16x3_amax')
    _lib.count_launch()
    return out


def _split_act(t, layout, amax):
    """fp32 [N,C,H,W] (any strides) -> the fp16 hi / lo operand layouts of csrc/conv_f16x3.cu; returns (tensor, Cp)."""
    lib = _lib.load()
    N, C, H, W = t.shape
    cp = _pad64(C)
    out = torch.empty([N, H, W, 3 * cp] if layout == 0 else [3 * N, H, W, cp], dtype=torch.float16, device=t.device)
    s = t.stride()
    _lib.check(lib.gt_f16x3_split_act(_lib.ptr(t), s[0], s[1], s[2], s[3], N, C, H, W, cp, layout, _lib.ptr(amax), _lib.ptr(out), _lib.stream_of(t)),
               'gt_f16x3_split_act')
    _lib.count_launch()
    return out, cp


def igemm_forward_f16x3(x, w, *, transpose, output_padding, stride, padding, groups):
    """fp32 convolution / transposed convolution on the fp16 tensor-core kernels (fp16 x 3).  Output: fp32, channels-last."""
    lib = _lib.load()
    N, cin, H, W = x.shape
    kh, kw = w.shape[2:]
    cout = w.shape[1] if transpose else w.shape[0]
    OH, OW = out_size(H, W, kh, kw, stride[0], padding[0], transpose)
    if OH <= 0 or OW <= 0:
        return None
    _log('fwd32', N, cin, cout, H, W, kh, stride[0], transpose)
    with torch.cuda.device(x.device):
        st = _lib.stream_of(x)
        ax, aw = _amax(x), _amax(w)
        xs, cinp = _split_act(x, 0, ax)
        coutp = _pad64(cout)
        sw = w.stride()
        s_co, s_ci = (sw[1], sw[0]) if transpose else (sw[0], sw[1])
        wp = torch.empty([kh * kw, coutp, 3 * cinp], dtype=torch.float16, device=x.device)
        _lib.check(lib.gt_f16x3_pack_weight(_lib.ptr(w), s_co, s_ci, sw[2], sw[3], cout, cin, kh, kw, coutp, cinp, _lib.ptr(aw), _lib.ptr(wp), st),
                   'gt_f16x3_pack_weight')
        # K-split: one partial sum per kernel row for stride-1 / strided kernels with several rows; the four parity phases of the
        # transposed stride-2 form are short already (<= 4 taps each)
        nslabs = kh if (kh > 1 and not (transpose and stride[0] == 2)) else 1
        slab = N * OH * OW * coutp
        part = torch.empty([nslabs, N, OH, OW, coutp], dtype=torch.float32, device=x.device)
        c3 = 3 * cinp
        _lib.check(lib.gt_conv2d_igemm_f16_f32out(_lib.ptr(xs), H * W * c3, W * c3, c3, _lib.ptr(wp), _lib.ptr(part), OH * OW * coutp, OW * coutp, coutp,
                                                  N, H, W, c3, OH, OW, coutp, kh, kw, stride[0], padding[0], 1 if transpose else 0,
                                                  slab if nslabs > 1 else 0, st), 'gt_conv2d_igemm_f16_f32out')
        y = torch.empty([N, OH, OW, cout], dtype=torch.float32, device=x.device)
        _lib.check(lib.gt_f16x3_slab_reduce(_lib.ptr(part), slab, nslabs, N * OH * OW, coutp, cout, _lib.ptr(ax), _lib.ptr(aw), _lib.ptr(y), st),
                   'gt_f16x3_slab_reduce')
        _lib.count_launch(3)
    return y.permute(0, 3, 1, 2)


def igemm_wgrad_f16x3(dy, x, weight_shape, *, transpose, output_padding, stride, padding, groups):
    """Weight gradient of an fp32 layer on the fp16 x 3 route: batch-concatenated hi / lo splits through the fp16 wgrad kernels."""
    kh, kw = weight_shape[2:]
    if not (enabled and f16x3_enabled) or dy.dtype != torch.float32 or x.dtype != torch.float32 or not x.is_cuda or groups != 1:
        return None
    if stride[0] != stride[1] or padding[0] != padding[1] or stride[0] not in (1, 2) or kh * kw > 9 or padding[0] >= 8:
        return None
    u, s = (x, dy) if transpose else (dy, x)
    if u.shape[1] != weight_shape[0] or s.shape[1] != weight_shape[1] or u.shape[1] < 16 or s.shape[1] < 16:
        return None
    if min(u.shape) == 0 or min(s.shape) == 0:
        return None
    lib = _lib.load()
    N, UC, UH, UW = u.shape
    _, SC, SH, SW = s.shape
    _log('wgrad32', N, SC, UC, SH, SW, kh, stride[0], transpose)
    with torch.cuda.device(x.device):
        au, as_ = _amax(u), _amax(s)
        us, ucp = _split_act(u, 1, au)
        ss, scp = _split_act(s, 2, as_)
        n3 = 3 * N
        nws = lib.gt_conv2d_wgrad_f16x3_workspace(n3, UH, UW, ucp, scp, kh, kw)
        ws = torch.empty([nws], dtype=torch.float32, device=x.device)
        dw = torch.empty(list(weight_shape), dtype=torch.float32, device=x.device)
        d = dw.stride()
        _lib.check(lib.gt_conv2d_wgrad_f16x3(_lib.ptr(us), UH * UW * ucp, UW * ucp, ucp, UH, UW, ucp, _lib.ptr(ss), SH * SW * scp, SW * scp, scp, SH, SW, scp,
                                             n3, kh, kw, stride[0], padding[0], _lib.ptr(dw), d[0], d[1], d[2], d[3], UC, SC, _lib.ptr(au), _lib.ptr(as_),
                                             _lib.ptr(ws), nws, _lib.stream_of(x)), 'gt_conv2d_wgrad_f16x3')
        _lib.count_launch(2)
    return dw


def igemm_forward(x, w, *, transpose, output_padding, stride, padding, groups, packed=None, epilogue=None):
    """epilogue: None, or (bias fp16 [Cout] or None, act code 1 / 3, alpha, gain, clamp[, addend]) -- the layer's bias_act (and an optional
    residual with the layout of the output, added after it) fused into the kernel's epilogue (fp16 only; the caller checks `covered` first)."""
    if epilogue is None and covered_fp32(x, w, transpose, output_padding, stride, padding, groups):
        return igemm_forward_f16x3(x, w, transpose=transpose, output_padding=output_padding, stride=stride, padding=padding, groups=groups)
    if not covered(x, w, transpose, output_padding, stride, padding, groups):
        return None
    lib = _lib.load()
    x = _nhwc(x)
    N, cin, H, W = x.shape
    kh, kw = w.shape[2:]
    cout = w.shape[1] if transpose else w.shape[0]
    OH, OW = out_size(H, W, kh, kw, stride[0], padding[0], transpose)
    if OH <= 0 or OW <= 0:
        return None
    _log('fwd', N, cin, cout, H, W, kh, stride[0], transpose)
    with torch.cuda.device(x.device):
        if packed is None:
            packed = pack_weight(w, transpose)
        y = torch.empty([N, cout, OH, OW], dtype=torch.float16, device=x.device, memory_format=torch.channels_last)
        ys_n, ys_h, ys_w = OH * OW * cout, OW * cout, cout
        if epilogue is None:
            _lib.check(lib.gt_conv2d_igemm_f16(_lib.ptr(x), H * W * cin, W * cin, cin, _lib.ptr(packed), _lib.ptr(y), ys_n, ys_h, ys_w,
                                               N, H, W, cin, OH, OW, cout, kh, kw, stride[0], padding[0], 1 if transpose else 0, _lib.stream_of(x)),
                       'gt_conv2d_igemm_f16')
        else:
            b, act, alpha, gain, clamp = epilogue[:5]
            addend = epilogue[5] if len(epilogue) > 5 else None
            if addend is None:
                _lib.check(lib.gt_conv2d_igemm_f16_bias_act(_lib.ptr(x), H * W * cin, W * cin, cin, _lib.ptr(packed), _lib.ptr(y), ys_n, ys_h, ys_w,
                                                            N, H, W, cin, OH, OW, cout, kh, kw, stride[0], padding[0], 1 if transpose else 0,
                                                            _lib.ptr(b), act, alpha, gain, clamp, _lib.stream_of(x)), 'gt_conv2d_igemm_f16_bias_act')
            else:
                assert addend.shape == y.shape and addend.dtype == torch.float16 and addend.stride() == y.stride(), 'the residual must have the layout of the output'
                _lib.check(lib.gt_conv2d_igemm_f16_bias_act_add(_lib.ptr(x), H * W * cin, W * cin, cin, _lib.ptr(packed), _lib.ptr(y), ys_n, ys_h, ys_w,
                                                                N, H, W, cin, OH, OW, cout, kh, kw, stride[0], padding[0], 1 if transpose else 0,
                                                                _lib.ptr(b), act, alpha, gain, clamp, _lib.ptr(addend), _lib.stream_of(x)),
                           'gt_conv2d_igemm_f16_bias_act_add')
        _lib.count_launch()
    return y


def igemm_wgrad(dy, x, weight_shape, *, transpose, output_padding, stride, padding, groups):
    kh, kw = weight_shape[2:]
    if dy.dtype == torch.float32 and x.dtype == torch.float32:
        return igemm_wgrad_f16x3(dy, x, weight_shape, transpose=transpose, output_padding=output_padding, stride=stride, padding=padding, groups=groups)
    if not enabled or dy.dtype != torch.float16 or x.dtype != torch.float16 or not x.is_cuda or groups != 1:
        return None
    if stride[0] != stride[1] or padding[0] != padding[1] or stride[0] not in (1, 2) or kh * kw > 9 or padding[0] >= 8:
        return None
    u, s = (x, dy) if transpose else (dy, x)
    if u.shape[1] != weight_shape[0] or s.shape[1] != weight_shape[1] or u.shape[1] % 64 != 0 or s.shape[1] % 64 != 0:
        return None
    if min(u.shape) == 0 or min(s.shape) == 0:
        return None
    lib = _lib.load()
    u, s = _nhwc(u), _nhwc(s)
    N, UC, UH, UW = u.shape
    _, SC, SH, SW = s.shape
    _log('wgrad', N, SC, UC, SH, SW, kh, stride[0], transpose)
    with torch.cuda.device(x.device):
        nws = lib.gt_conv2d_wgrad_workspace(N, UH, UW, UC, SC, kh, kw)
        ws = torch.empty([nws], dtype=torch.float32, device=x.device)
        dw = torch.empty(list(weight_shape), dtype=torch.float16, device=x.device)
        d = dw.stride()
        _lib.check(lib.gt_conv2d_wgrad_f16(_lib.ptr(u), UH * UW * UC, UW * UC, UC, UH, UW, UC, _lib.ptr(s), SH * SW * SC, SW * SC, SC, SH, SW, SC,
                                           N, kh, kw, stride[0], padding[0], _lib.ptr(dw), d[0], d[1], d[2], d[3], _lib.ptr(ws), nws, _lib.stream_of(x)),
                   'gt_conv2d_wgrad_f16')
        _lib.count_launch(2)
    return dw


def install():
    from . import conv_backend
    conv_backend._igemm_forward = igemm_forward
    conv_backend._igemm_wgrad = igemm_wgrad
