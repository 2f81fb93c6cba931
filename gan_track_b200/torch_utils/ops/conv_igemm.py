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


# fp32 layers on the tensor cores as 3 x TF32.  OFF by default: measured on B200 (tools/test_tf32x3.py) the products are
# fp32-accurate but the tensor core adds every K = 8 MMA into its fp32 accumulator with truncation, and that bias grows
# linearly with the number of accumulator updates -- 7e-6 relative for the 64..128-channel 3x3 cases (<= 432 updates) but
# 3e-5 .. 5e-5 for the 512-channel 3x3 layers of the fp32 blocks (1728 updates), above the 1e-5 the north star asks of fp32
# layers (cuDNN's own fp32 kernels sit at 2e-5 there).  Needs two-level accumulation (TMEM partials promoted to fp32
# registers every ~64 updates) before it can take over; until then those layers keep the library's true-fp32 route.
tf32x3_enabled = False


def covered_fp32(x, w, transpose, output_padding, stride, padding, groups):
    if not (enabled and tf32x3_enabled) or x.dtype != torch.float32 or w.dtype != torch.float32 or not x.is_cuda or groups != 1:
        return False
    if stride[0] != stride[1] or padding[0] != padding[1] or tuple(output_padding) != (0, 0) or stride[0] not in (1, 2):
        return False
    kh, kw = w.shape[2:]
    cin, cout = (w.shape[0], w.shape[1]) if transpose else (w.shape[1], w.shape[0])
    if x.shape[1] != cin or cin % 32 != 0 or cout % 64 != 0 or kh * kw > 9 or padding[0] >= 8:
        return False
    if transpose and stride[0] == 2 and padding[0] != 0:
        return False
    return min(x.shape) > 0


def igemm_forward_tf32x3(x, w, *, transpose, output_padding, stride, padding, groups):
    """fp32 convolution / transposed convolution as ONE TF32 implicit GEMM over [x_big | x_big | x_small] x
    [w_big | w_small | w_big] (csrc/conv_igemm.cu).  Output: fp32, channels-last."""
    lib = _lib.load()
    N, cin, H, W = x.shape
    kh, kw = w.shape[2:]
    cout = w.shape[1] if transpose else w.shape[0]
    OH, OW = out_size(H, W, kh, kw, stride[0], padding[0], transpose)
    if OH <= 0 or OW <= 0:
        return None
    with torch.cuda.device(x.device):
        st = _lib.stream_of(x)
        xs = torch.empty([N, H, W, 3 * cin], dtype=torch.float32, device=x.device)
        sx = x.stride()
        _lib.check(lib.gt_split_tf32x3(_lib.ptr(x), sx[0], sx[1], sx[2], sx[3], N, cin, H, W, _lib.ptr(xs), st), 'gt_split_tf32x3')
        sw = w.stride()
        s_co, s_ci = (sw[1], sw[0]) if transpose else (sw[0], sw[1])
        wp = torch.empty([kh * kw, cout, 3 * cin], dtype=torch.float32, device=x.device)
        _lib.check(lib.gt_conv_pack_weight_tf32x3(_lib.ptr(w), s_co, s_ci, sw[2], sw[3], cout, cin, kh, kw, _lib.ptr(wp), st), 'gt_conv_pack_weight_tf32x3')
        y = torch.empty([N, cout, OH, OW], dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        if y.stride(1) != 1 or y.stride(3) != cout:            # size-1 dims can leave ambiguous strides; force NHWC
            y = torch.empty([N, OH, OW, cout], dtype=torch.float32, device=x.device).permute(0, 3, 1, 2)
        c3 = 3 * cin
        _lib.check(lib.gt_conv2d_igemm_tf32(_lib.ptr(xs), H * W * c3, W * c3, c3, _lib.ptr(wp), _lib.ptr(y), OH * OW * cout, OW * cout, cout,
                                            N, H, W, c3, OH, OW, cout, kh, kw, stride[0], padding[0], 1 if transpose else 0, st),
                   'gt_conv2d_igemm_tf32')
        _lib.count_launch(3)
    return y


def igemm_forward(x, w, *, transpose, output_padding, stride, padding, groups, packed=None, epilogue=None):
    """epilogue: None, or (bias fp16 [Cout] or None, act code 1 / 3, alpha, gain, clamp) -- the layer's bias_act fused into
    the kernel's epilogue (fp16 only; the caller checks `covered` first)."""
    if epilogue is None and covered_fp32(x, w, transpose, output_padding, stride, padding, groups):
        return igemm_forward_tf32x3(x, w, transpose=transpose, output_padding=output_padding, stride=stride, padding=padding, groups=groups)
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
            b, act, alpha, gain, clamp = epilogue
            _lib.check(lib.gt_conv2d_igemm_f16_bias_act(_lib.ptr(x), H * W * cin, W * cin, cin, _lib.ptr(packed), _lib.ptr(y), ys_n, ys_h, ys_w,
                                                        N, H, W, cin, OH, OW, cout, kh, kw, stride[0], padding[0], 1 if transpose else 0,
                                                        _lib.ptr(b), act, alpha, gain, clamp, _lib.stream_of(x)), 'gt_conv2d_igemm_f16_bias_act')
        _lib.count_launch()
    return y


def igemm_wgrad(dy, x, weight_shape, *, transpose, output_padding, stride, padding, groups):
    kh, kw = weight_shape[2:]
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
