"""G_ema sampling throughput (the snapshot / metrics / gen_images path): images per second of
`G_ema(z, c, noise_mode='const')` in eval mode at the CLARO 256x256 configuration, eager and as one CUDA graph.
python tools/bench_sampling.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_track_b200.training import training_loop as tl  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda', 0)
cfg = tl.claro_config(resolution=256, batch=B, aug='ada')
tr = tl.Trainer(cfg, device=dev, use_graphs=False)
G = tr.G_ema.eval().requires_grad_(False)
z = torch.randn([B, 512], device=dev)
c = torch.nn.functional.one_hot(torch.randint(0, 2, [B]), 2).float().to(dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters


with torch.no_grad():
    ms = timeit(lambda: G(z, c, noise_mode='const'))
    print(f'eager : batch {B}: {ms:7.2f} ms  = {B / ms * 1e3:8.0f} img/s', flush=True)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        G(z, c, noise_mode='const')
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        out = G(z, c, noise_mode='const')
    ms = timeit(g.replay)
    print(f'graph : batch {B}: {ms:7.2f} ms  = {B / ms * 1e3:8.0f} img/s   (29.77 GFLOP/img -> {29.77e9 * B / ms / 1e9:6.0f} TFLOP/s)', flush=True)
    assert torch.isfinite(out).all()
