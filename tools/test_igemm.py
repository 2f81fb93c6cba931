"""GPU check + timing of the tcgen05 implicit-GEMM convolution against torch's own convolution (fp32 math on the same
fp16 inputs for the check; cuDNN fp16 channels-last for the timing).  Run on the GPU box:
    timeout 600 python tools/test_igemm.py [--time]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from gan_track_b200.torch_utils.ops import conv_igemm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--time', action='store_true')
ap.add_argument('--wgrad', action='store_true')
ap.add_argument('--variant', type=int, default=0)
args = ap.parse_args()
dev = torch.device('cuda', 0)
from gan_track_b200 import _lib  # noqa: E402
_lib.load().gt_conv_igemm_config(args.variant)
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True

# (name, N, Cin, Cout, H, W, k, stride, pad, transpose)
CASES = [
    ('3x3 p1 64->64 32x32', 2, 64, 64, 32, 32, 3, 1, 1, False),
    ('3x3 p1 128->256 16x16', 4, 128, 256, 16, 16, 3, 1, 1, False),
    ('3x3 p1 64->128 33x33', 3, 64, 128, 33, 33, 3, 1, 1, False),
    ('3x3 p1 64->64 8x8', 8, 64, 64, 8, 8, 3, 1, 1, False),
    ('3x3 p1 64->64 4x4', 5, 64, 64, 4, 4, 3, 1, 1, False),
    ('3x3 p1 512->512 32x32', 2, 512, 512, 32, 32, 3, 1, 1, False),
    ('1x1 128->64 32x32', 2, 128, 64, 32, 32, 1, 1, 0, False),
    ('3x3 s2 64->128 33x33', 2, 64, 128, 33, 33, 3, 2, 0, False),
    ('3x3 s2 64->64 257x257', 2, 64, 64, 257, 257, 3, 2, 0, False),
    ('3x3 T s2 128->64 16x16', 2, 128, 64, 16, 16, 3, 2, 0, True),
    ('3x3 T s2 64->64 128x128', 1, 64, 64, 128, 128, 3, 2, 0, True),
    ('3x3 T s1 p1 64->128 32x32', 2, 64, 128, 32, 32, 3, 1, 1, True),
    ('3x3 T s1 p0 64->64 16x16', 2, 64, 64, 16, 16, 3, 1, 0, True),
    ('3x3 p1 64->64 64x64', 2, 64, 64, 64, 64, 3, 1, 1, False),
    ('3x3 p1 128->128 40x72', 3, 128, 128, 40, 72, 3, 1, 1, False),
    ('3x3 p1 256->512 32x32', 2, 256, 512, 32, 32, 3, 1, 1, False),
    ('1x1 128->256 64x64', 2, 128, 256, 64, 64, 1, 1, 0, False),
    ('3x3 T s2 128->64 32x32', 2, 128, 64, 32, 32, 3, 2, 0, True),
    ('3x3 T s2 256->128 64x64', 1, 256, 128, 64, 64, 3, 2, 0, True),
    ('3x3 p1 64->64 128x128 n8', 8, 64, 64, 128, 128, 3, 1, 1, False),
    ('3x3 p1 128->64 72x64 n4', 4, 128, 64, 72, 64, 3, 1, 1, False),
    ('3x3 T s1 p1 64->128 40x40 n6', 6, 64, 128, 40, 40, 3, 1, 1, True),
    ('3x3 p0 64->64 66x66 n4', 4, 64, 64, 66, 66, 3, 1, 0, False),
    ('3x3 p1 256->256 32x32 n8', 8, 256, 256, 32, 32, 3, 1, 1, False),
    # row-streaming kernel (csrc/conv_rows.cu): 64 -> 64, width >= 128; ragged widths / heights, transposed (= data gradient) form
    ('3x3 p1 64->64 256x256 n2', 2, 64, 64, 256, 256, 3, 1, 1, False),
    ('3x3 T s1 p1 64->64 130x200 n3', 3, 64, 64, 130, 200, 3, 1, 1, True),
    ('3x3 p1 64->64 9x300 n2', 2, 64, 64, 9, 300, 3, 1, 1, False),
    ('3x3 p1 64->64 257x129 n1', 1, 64, 64, 257, 129, 3, 1, 1, False),
]


def ref_conv(x, w, stride, pad, transpose):
    xf, wf = x.float(), w.float()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        if transpose:
            return F.conv_transpose2d(xf, wf, stride=stride, padding=pad)
        return F.conv2d(xf, wf, stride=stride, padding=pad)
    finally:
        torch.backends.cudnn.allow_tf32 = old


fails = 0
for name, N, ci, co, H, W, k, s, p, tr in CASES:
    x = torch.randn([N, ci, H, W], device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
    wshape = [ci, co, k, k] if tr else [co, ci, k, k]
    w = (torch.randn(wshape, device=dev) / (ci * k * k) ** 0.5).to(torch.float16)
    y = conv_igemm.igemm_forward(x, w, transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
    torch.cuda.synchronize()
    assert y is not None, name
    r = ref_conv(x, w, s, p, tr)
    err = float((y.float() - r).abs().max() / r.abs().max())
    ok = err < 2e-3 and y.shape == r.shape
    fails += not ok
    print(f'{"ok  " if ok else "FAIL"} {name:32s} out {tuple(y.shape)} rel err {err:.2e}', flush=True)
    if not ok and y.shape == r.shape:
        d = (y.float() - r).abs()
        idx = torch.nonzero(d > 2e-3 * r.abs().max())
        print('   mismatches:', idx.shape[0], 'first:', idx[:5].tolist(), flush=True)

if args.wgrad:
    for name, N, ci, co, H, W, k, s, p, tr in CASES:
        x = torch.randn([N, ci, H, W], device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
        wshape = [ci, co, k, k] if tr else [co, ci, k, k]
        OH, OW = conv_igemm.out_size(H, W, k, k, s, p, tr)
        dy = torch.randn([N, co, OH, OW], device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
        dw = conv_igemm.igemm_wgrad(dy, x, tuple(wshape), transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
        torch.cuda.synchronize()
        if dw is None:
            print(f'skip {name:32s} wgrad not covered', flush=True)
            continue
        wd = torch.zeros(wshape, device=dev)
        if tr:
            _, r, _ = torch.ops.aten.convolution_backward(dy.float(), x.float(), wd, None, [s, s], [p, p], [1, 1], True, [0, 0], 1, [False, True, False])
        else:
            _, r, _ = torch.ops.aten.convolution_backward(dy.float(), x.float(), wd, None, [s, s], [p, p], [1, 1], False, [0, 0], 1, [False, True, False])
        err = float((dw.float() - r).abs().max() / r.abs().max())
        ok = err < 2e-3 and dw.shape == r.shape
        fails += not ok
        print(f'{"ok  " if ok else "FAIL"} wgrad {name:32s} rel err {err:.2e}', flush=True)

print('FAILURES:', fails, flush=True)

if args.time and fails == 0:
    # the fp16 layers of the 256x256 cbase-16384 networks at batch 32
    T = [
        ('G b32 conv1 512->512 3x3 @32', 32, 512, 512, 32, 32, 3, 1, 1, False),
        ('G b64 conv0 512->256 T s2 @32->65', 32, 512, 256, 32, 32, 3, 2, 0, True),
        ('G b64 conv1 256->256 3x3 @64', 32, 256, 256, 64, 64, 3, 1, 1, False),
        ('G b128 conv0 256->128 T s2 @64->129', 32, 256, 128, 64, 64, 3, 2, 0, True),
        ('G b128 conv1 128->128 3x3 @128', 32, 128, 128, 128, 128, 3, 1, 1, False),
        ('G b256 conv0 128->64 T s2 @128->257', 32, 128, 64, 128, 128, 3, 2, 0, True),
        ('G b256 conv1 64->64 3x3 @256', 32, 64, 64, 256, 256, 3, 1, 1, False),
        ('D b256 conv1 64->128 s2 @257', 32, 64, 128, 257, 257, 3, 2, 0, False),
        ('D b128 conv0 128->128 3x3 @128', 32, 128, 128, 128, 128, 3, 1, 1, False),
        ('D b64 conv1 256->512 s2 @65', 32, 256, 512, 65, 65, 3, 2, 0, False),
        ('D b128 skip 128->256 1x1 @64', 32, 128, 256, 64, 64, 1, 1, 0, False),
    ]
    flush = torch.empty([256 << 20], dtype=torch.uint8, device=dev)

    def bench(fn, iters=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            fn()
            e_.record()
            e_.synchronize()
            ts.append(s_.elapsed_time(e_))
        return sum(ts) / len(ts)

    for name, N, ci, co, H, W, k, s, p, tr in T:
        x = torch.randn([N, ci, H, W], device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
        wshape = [ci, co, k, k] if tr else [co, ci, k, k]
        w = (torch.randn(wshape, device=dev) / (ci * k * k) ** 0.5).to(torch.float16).contiguous(memory_format=torch.channels_last)
        OH, OW = conv_igemm.out_size(H, W, k, k, s, p, tr)
        flops = 2.0 * N * ci * co * k * k * (H * W if tr else OH * OW)
        pk = conv_igemm.pack_weight(w, tr)
        kw_ = dict(transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
        t_full = bench(lambda: conv_igemm.igemm_forward(x, w, **kw_))
        res = {}
        for v in (0, 2, 3, 1, 7):
            _lib.load().gt_conv_igemm_config(v)
            res[v] = bench(lambda: conv_igemm.igemm_forward(x, w, packed=pk, **kw_))
        _lib.load().gt_conv_igemm_config(args.variant)
        t_ours, t_v2, t_v3, t_v1, t_v5 = res[0], res[2], res[3], res[1], res[7]
        if tr:
            t_lib = bench(lambda: F.conv_transpose2d(x, w, stride=s, padding=p))
        else:
            t_lib = bench(lambda: F.conv2d(x, w, stride=s, padding=p))
        line = f'{name:40s} fwd ours {t_ours * 1e3:7.1f} us {flops / t_ours / 1e9:6.0f} TF/s (+pack {t_full * 1e3:6.1f}; per-tap {t_v1 * 1e3:6.1f}, cfg2 {t_v2 * 1e3:6.1f}, cfg3 {t_v3 * 1e3:6.1f}, single-CTA {t_v5 * 1e3:6.1f}) | cudnn {t_lib * 1e3:7.1f} us {flops / t_lib / 1e9:6.0f} TF/s'
        if args.wgrad:
            dy = torch.randn([N, co, OH, OW], device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
            t_w = bench(lambda: conv_igemm.igemm_wgrad(dy, x, tuple(wshape), transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1))
            _lib.load().gt_conv_wgrad_config(1)
            t_w1 = bench(lambda: conv_igemm.igemm_wgrad(dy, x, tuple(wshape), transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1))
            _lib.load().gt_conv_wgrad_config(0)
            t_wl = bench(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [s, s], [p, p], [1, 1], tr, [0, 0], 1, [False, True, False]))
            line += f' || wgrad ours {t_w * 1e3:7.1f} us {flops / t_w / 1e9:6.0f} TF/s (per-tap-row {t_w1 * 1e3:6.1f}) | cudnn {t_wl * 1e3:7.1f} us {flops / t_wl / 1e9:6.0f} TF/s'
        print(line, flush=True)
sys.exit(1 if fails else 0)
