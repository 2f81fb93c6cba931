"""Microbenchmark of the small-batch fully-connected kernels (csrc/fc.cu) against torch's addmm.  python tools/bench_fc.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_track_b200.torch_utils.ops import fc  # noqa: E402

dev = torch.device('cuda', 0)
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters * 1e3


for M, I, O in [(32, 512, 512), (32, 1024, 512), (32, 8192, 512), (32, 512, 64), (16, 512, 512)]:
    x = torch.randn(M, I, device=dev)
    w = torch.randn(O, I, device=dev)
    b = torch.randn(O, device=dev)
    dy = torch.randn(M, O, device=dev)
    g = torch.cuda.CUDAGraph()
    res = {}
    for name, fn in [('fwd ours', lambda: fc._fwd(x, w, b, 0.1, 1.0)), ('fwd addmm', lambda: torch.addmm(b.unsqueeze(0), x, (w * 0.1).t())),
                     ('dgrad ours', lambda: fc._dgrad(dy, w, 0.1)), ('dgrad mm', lambda: dy.matmul(w * 0.1)),
                     ('wgrad ours', lambda: fc._wgrad(dy, x, 0.1, 1.0, True)), ('wgrad mm', lambda: (dy.t().matmul(x) * 0.1, dy.sum(0)))]:
        # time inside a CUDA graph (what the training step does): 20 launches per replay
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                fn()
        res[name] = timeit(gr.replay, iters=20) / 20
    print(f'M={M} I={I} O={O}: ' + '  '.join(f'{k} {v:6.1f} us' for k, v in res.items()), flush=True)
