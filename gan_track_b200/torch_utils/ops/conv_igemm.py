"""Binding of the tcgen05 implicit-GEMM convolution (csrc/conv_igemm.cu, csrc/conv_wgrad.cu) behind `conv_backend`.

`igemm_forward` / `igemm_wgrad` return None when a call is outside the kernels' coverage, in which case
`conv_backend` takes the library route the reference itself uses.  Coverage: fp16, groups = 1, Cin and Cout multiples
of 64, kernels of at most 9 taps, stride 1 (any pad), stride 2, transposed stride 1, transposed stride 2 with pad 0 and
no output padding.  Tensors are consumed channels-last (NHWC); an NCHW input is re-laid-out first.
"""
import torch

from ... import _lib

enabled = True


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


def igemm_forward(x, w, *, transpose, output_padding, stride, padding, groups, packed=None):
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
    with torch.cuda.device(x.device):
        if packed is None:
            packed = pack_weight(w, transpose)
        y = torch.empty([N, cout, OH, OW], dtype=torch.float16, device=x.device, memory_format=torch.channels_last)
        ys_n, ys_h, ys_w = OH * OW * cout, OW * cout, cout
        _lib.check(lib.gt_conv2d_igemm_f16(_lib.ptr(x), H * W * cin, W * cin, cin, _lib.ptr(packed), _lib.ptr(y), ys_n, ys_h, ys_w,
                                           N, H, W, cin, OH, OW, cout, kh, kw, stride[0], padding[0], 1 if transpose else 0, _lib.stream_of(x)),
                   'gt_conv2d_igemm_f16')
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
