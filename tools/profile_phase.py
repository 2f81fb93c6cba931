"""Kernel-time breakdown of ONE phase (Gmain / Greg / Dmain / Dreg) run eagerly.  python tools/profile_phase.py Greg"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from gan_track_b200.training import training_loop as tl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'Greg'
out = sys.argv[2] if len(sys.argv) > 2 else f'gpurun_out/phase_{name}.txt'
dev = torch.device('cuda', 0)
B = 32
cfg = tl.claro_config(resolution=256, batch=B, aug='ada')
tr = tl.Trainer(cfg, device=dev, use_graphs=False)
img = [(torch.rand([B, 1, 256, 256], device=dev) * 2 - 1)]
c = [torch.nn.functional.one_hot(torch.randint(0, 2, [B]), 2).float().to(dev)]
z = [torch.randn([B, 512], device=dev)]
phase = [p for p in tr.phases if p.name == name][0]
for _ in range(2):
    tr._run_phase_eager(phase, img, c, z, c)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr._run_phase_eager(phase, img, c, z, c)
    torch.cuda.synchronize()
rows = sorted([r for r in prof.key_averages() if r.device_time_total > 0], key=lambda r: -r.device_time_total)
total = sum(r.device_time_total for r in rows)
with open(out, 'w') as f:
    f.write(f'phase {name}: gpu busy {total / 1e3:.2f} ms, kernels {sum(r.count for r in rows)}\n')
    for r in rows[:int(os.environ.get("GT_PROFILE_ROWS", "70"))]:
        f.write(f'{r.device_time_total / total * 100:6.2f} {r.device_time_total / 1e3:9.3f} ms {r.count:6d} x {r.device_time_total / max(r.count, 1):9.1f} us  {r.key[:140]}\n')
print(open(out).read())
