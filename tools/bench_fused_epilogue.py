"""A/B per layer: bias_act in the convolution epilogue (conv2d_gradfix.fuse_bias_act) vs convolution + separate bias_act pass, at the
discriminator's fp16 layer shapes (batch 64 = the merged Dmain pass).  python tools/bench_fused_epilogue.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gan_track_b200.torch_utils.ops import conv2d_gradfix  # noqa: E402

dev = torch.device('cuda', 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
# (name, Cin, Cout, H, k, stride, pad, act, gain)
CASES = [('b256 conv0 64->64 3x3', 64, 64, 256, 3, 1, 1, 'lrelu', np.sqrt(2)), ('b256 conv1 64->128 s2 (blurred 257)', 64, 128, 257, 3, 2, 0, 'lrelu', 1.0),
         ('b256 skip 64->128 1x1 @128', 64, 128, 128, 1, 1, 0, 'linear', np.sqrt(0.5)),
         ('b128 conv0 128->128 3x3', 128, 128, 128, 3, 1, 1, 'lrelu', np.sqrt(2)), ('b128 conv1 128->256 s2', 128, 256, 129, 3, 2, 0, 'lrelu', 1.0),
         ('b64 conv0 256->256 3x3', 256, 256, 64, 3, 1, 1, 'lrelu', np.sqrt(2)), ('b64 conv1 256->512 s2', 256, 512, 65, 3, 2, 0, 'lrelu', 1.0),
         ('b32 conv0 512->512 3x3', 512, 512, 32, 3, 1, 1, 'lrelu', np.sqrt(2)), ('b32 conv1 512->512 s2', 512, 512, 33, 3, 2, 0, 'lrelu', 1.0)]


def timeit(fn, xs, iters=20):
    for x in xs:
        fn(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(xs[i % len(xs)])
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters * 1e3


print(f'# batch {N}; us per layer forward (rotating inputs), separate = convolution + bias_act pass, fused = bias_act in the convolution epilogue')
for name, ci, co, H, k, st, pad, act, gain in CASES:
    nset = max(2, min(6, int(300e6 // (N * ci * H * H * 2)) + 1))
    xs = [torch.randn([N, ci, H, H], device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last) for _ in range(nset)]
    w = (torch.randn([co, ci, k, k], device=dev) / np.sqrt(ci * k * k)).to(torch.float16)
    b = torch.randn([co], device=dev).to(torch.float16) if act == 'lrelu' else None
    res = {}
    for fused in (False, True):
        conv2d_gradfix.fuse_bias_act = fused
        with torch.no_grad():
            res[fused] = timeit(lambda x: conv2d_gradfix.conv2d_bias_act(x, w, b, act=act, gain=float(gain), clamp=256.0, stride=st, padding=pad), xs)
    conv2d_gradfix.fuse_bias_act = False
    print(f'  {name:38s} separate {res[False]:8.1f}   fused {res[True]:8.1f}   saved {res[False] - res[True]:7.1f}')
    del xs
