"""HBM-roofline microbenchmark of the memory-bound kernels at their training shapes (batch 32, 256x256 config):
achieved GB/s = algorithmic bytes / CUDA-event time, L2 flushed between launches.  python tools/bench_hbm_ops.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_track_b200.torch_utils.ops import bias_act, upfirdn2d  # noqa: E402

dev = torch.device('cuda', 0)
peak = 6476.4
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.exists(pk):
    peak = json.load(open(pk))['hbm_gbs']
flush = torch.empty([256 << 20], dtype=torch.uint8, device=dev)


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)


f = upfirdn2d.setup_filter([1, 3, 3, 1], device=dev)
rows = []
for name, shape, dt, cl, kw in [
    ('blur after up-conv  [32,64,257,257] f16 NHWC', (32, 64, 257, 257), torch.float16, True, dict(padding=[1, 1, 1, 1], gain=4)),
    ('blur after up-conv  [32,128,129,129] f16 NHWC', (32, 128, 129, 129), torch.float16, True, dict(padding=[1, 1, 1, 1], gain=4)),
    ('blur after up-conv  [32,512,33,33] f16 NHWC', (32, 512, 33, 33), torch.float16, True, dict(padding=[1, 1, 1, 1], gain=4)),
    ('blur before down-conv [32,64,256,256] f16 NHWC', (32, 64, 256, 256), torch.float16, True, dict(padding=[2, 2, 2, 2])),
    ('skip downsample     [32,64,256,256] f16 NHWC', (32, 64, 256, 256), torch.float16, True, dict(down=2, padding=[1, 1, 1, 1])),
    ('skip downsample     [32,256,64,64] f16 NHWC', (32, 256, 64, 64), torch.float16, True, dict(down=2, padding=[1, 1, 1, 1])),
    ('skip downsample bwd [32,64,128,128] f16 NHWC', (32, 64, 128, 128), torch.float16, True, dict(up=2, padding=[2, 1, 2, 1], gain=4)),
    ('skip downsample bwd [32,256,32,32] f16 NHWC', (32, 256, 32, 32), torch.float16, True, dict(up=2, padding=[2, 1, 2, 1], gain=4)),
    ('blur fp32 planes    [32,512,33,33] f32 NCHW', (32, 512, 33, 33), torch.float32, False, dict(padding=[1, 1, 1, 1], gain=4)),
    ('blur fp32 planes    [32,512,17,17] f32 NCHW', (32, 512, 17, 17), torch.float32, False, dict(padding=[1, 1, 1, 1], gain=4)),
    ('img upsample        [32,1,128,128] f32 NCHW', (32, 1, 128, 128), torch.float32, False, dict(up=2, padding=[2, 1, 2, 1], gain=4)),
]:
    x = torch.randn(shape, device=dev).to(dt)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    y = upfirdn2d.upfirdn2d(x, f, **kw)
    nbytes = (x.numel() + y.numel()) * x.element_size()
    ms = bench(lambda: upfirdn2d.upfirdn2d(x, f, **kw))
    rows.append((f'upfirdn2d {name}', nbytes, ms))

from gan_track_b200 import _lib  # noqa: E402
lib = _lib.load()
for variant, vname in [(0, 'bulk-staged'), (1, 'direct')]:
    lib.gt_stream_config(variant)
    for shape in [(32, 64, 256, 256), (32, 256, 64, 64)]:
        xb = torch.randn(shape, device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
        b = torch.randn([shape[1]], device=dev, dtype=torch.float16)
        ms = bench(lambda: bias_act.bias_act(xb, b, act='lrelu', clamp=256.0))
        rows.append((f'bias_act fwd lrelu {list(shape)} f16 NHWC ({vname})', 2 * xb.numel() * 2, ms))
        xg = xb.clone().requires_grad_(True)
        bg = b.clone().requires_grad_(True)
        yb = bias_act.bias_act(xg, bg, act='lrelu', clamp=256.0)
        dy = torch.randn_like(yb)
        ms = bench(lambda: torch.autograd.grad(yb, [xg, bg], dy, retain_graph=True))
        rows.append((f'bias_act bwd (dx + db fused) {list(shape)} f16 NHWC ({vname})', 3 * xb.numel() * 2, ms))
        del xb, xg, yb, dy
lib.gt_stream_config(0)
# reference point: plain device-to-device copy of the same tensor size
xc = torch.randn([32, 64, 256, 256], device=dev).to(torch.float16)
yc = torch.empty_like(xc)
ms = bench(lambda: yc.copy_(xc))
rows.append(('torch copy_ [32,64,256,256] f16 (reference point)', 2 * xc.numel() * 2, ms))

# modulation / demodulation+activation element-wise halves of the modulated convolution
from gan_track_b200.torch_utils.ops import modulated  # noqa: E402
for shape in [(32, 64, 256, 256), (32, 128, 128, 128)]:
    n, c, h, w = shape
    xm = torch.randn(shape, device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    sm = torch.randn([n, c], device=dev).requires_grad_(True)
    dm = torch.rand([n, c], device=dev).requires_grad_(True)
    nz = torch.randn([n, 1, h, w], device=dev).to(torch.float16).requires_grad_(True)
    bm = torch.randn([c], device=dev).to(torch.float16).requires_grad_(True)
    nb = xm.numel() * 2
    ms = bench(lambda: modulated.mod_scale(xm.detach(), sm.detach()))
    rows.append((f'mod_scale fwd {list(shape)} f16 NHWC', 2 * nb, ms))
    ym = modulated.mod_scale(xm, sm)
    gy = torch.randn_like(ym)
    ms = bench(lambda: torch.autograd.grad(ym, [xm, sm], gy, retain_graph=True))
    rows.append((f'mod_scale bwd (gx + gs) {list(shape)} f16 NHWC', 3 * nb, ms))
    ms = bench(lambda: modulated.demod_act(xm.detach(), dm.detach(), nz.detach(), bm.detach(), act='lrelu', gain=1.4142, clamp=256.0))
    rows.append((f'demod_act fwd {list(shape)} f16 NHWC', 2 * nb, ms))
    yd = modulated.demod_act(xm, dm, nz, bm, act='lrelu', gain=1.4142, clamp=256.0)
    ms = bench(lambda: torch.autograd.grad(yd, [xm, dm, nz, bm], gy, retain_graph=True))
    rows.append((f'demod_act bwd (gx + gd + gnoise + db) {list(shape)} f16 NHWC', 4 * nb, ms))
    del xm, ym, yd, gy
for name, nbytes, ms in rows:
    gbs = nbytes / ms / 1e6
    print(f'{name:72s} {ms * 1e3:8.1f} us {gbs:8.0f} GB/s  {gbs / peak:5.2f} of measured HBM peak ({peak:.0f} GB/s)', flush=True)
