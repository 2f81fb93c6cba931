"""Kernel-time breakdown of training iterations (torch.profiler, CUDA activity only).  Usage on the GPU box:
    python tools/profile_step.py [--steps 16] [--out gpurun_out/step_profile.txt]
Prints total GPU-busy time per kernel name and the wall time per step, so launch-bound vs GPU-bound is visible."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from gan_track_b200.training import training_loop as tl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=16)
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--res', type=int, default=256)
ap.add_argument('--out', default='gpurun_out/step_profile.txt')
ap.add_argument('--aug', default='ada')
ap.add_argument('--graphs', action='store_true')
a = ap.parse_args()

dev = torch.device('cuda', 0)
cfg = tl.claro_config(resolution=a.res, batch=a.batch, aug=a.aug)
tr = tl.Trainer(cfg, device=dev, use_graphs=a.graphs)
img = torch.rand([a.batch, 1, a.res, a.res], device=dev) * 255
c = torch.nn.functional.one_hot(torch.randint(0, 2, [a.batch]), 2).float().to(dev)
for _ in range(17 if a.graphs else 3):
    tr.train_step(img, c)
torch.cuda.synchronize()
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        tr.train_step(img, c)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
total = sum(r.device_time_total for r in rows)
os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
with open(a.out, 'w') as f:
    f.write(f'steps={a.steps} batch={a.batch} res={a.res} wall_ms_per_step={wall / a.steps * 1e3:.2f} gpu_busy_ms_per_step={total / a.steps / 1e3:.2f} '
            f'kernels_per_step={sum(r.count for r in rows) / a.steps:.0f}\n')
    f.write(f'{"share":>6} {"ms/step":>9} {"calls/step":>10} {"us/call":>9}  name\n')
    small = [r for r in rows if r.device_time_total / max(r.count, 1) < 6.0]
    f.write(f'# kernels averaging < 6 us: {sum(r.count for r in small) / a.steps:.0f} launches/step, {sum(r.device_time_total for r in small) / a.steps / 1e3:.2f} ms/step\n')
    for r in rows[:110]:
        f.write(f'{r.device_time_total / total * 100:6.2f} {r.device_time_total / a.steps / 1e3:9.3f} {r.count / a.steps:10.1f} '
                f'{r.device_time_total / max(r.count, 1):9.1f}  {r.key[:150]}\n')
print(open(a.out).read()[:6000])
