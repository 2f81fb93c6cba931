"""Where do the small kernels come from?  Kernel launch counts and GPU time per component of one training iteration
(torch.profiler, CUDA activity), each component run on its own.  python tools/count_launches.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from gan_track_b200.training import training_loop as tl  # noqa: E402

dev = torch.device('cuda', 0)
B = 32
cfg = tl.claro_config(resolution=256, batch=B, aug='ada')
tr = tl.Trainer(cfg, device=dev, use_graphs=False)
G, D, aug = tr.G, tr.D, tr.augment_pipe
aug.p.fill_(0.3)
img = torch.rand([B, 1, 256, 256], device=dev) * 2 - 1
c = torch.nn.functional.one_hot(torch.randint(0, 2, [B]), 2).float().to(dev)
z = torch.randn([B, 512], device=dev)


def measure(name, fn, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    rows = prof.key_averages()
    n = sum(r.count for r in rows)
    t = sum(r.device_time_total for r in rows)
    small = [r for r in rows if r.device_time_total / max(r.count, 1) < 6.0]
    ns = sum(r.count for r in small)
    ts = sum(r.device_time_total for r in small)
    print(f'{name:44s} kernels {n:6d}  gpu {t / 1e3:8.3f} ms   | < 6 us: {ns:6d} kernels {ts / 1e3:7.3f} ms', flush=True)
    top = sorted(small, key=lambda r: -r.count)[:6]
    for r in top:
        print(f'      {r.count:5d} x {r.device_time_total / max(r.count, 1):5.1f} us  {r.key[:110]}')


def set_grad(m, flag):
    m.requires_grad_(flag)


def g_map_fwd_bwd():
    set_grad(G, True)
    ws = G.mapping(z, c)
    ws.sum().backward()
    set_grad(G, False)
    G.zero_grad(set_to_none=True)


ws_fixed = G.mapping(z, c).detach()


def g_syn_fwd():
    with torch.no_grad():
        G.synthesis(ws_fixed)


def g_syn_fwd_bwd():
    set_grad(G, True)
    w = ws_fixed.clone().requires_grad_(True)
    out = G.synthesis(w)
    out.sum().backward()
    set_grad(G, False)
    G.zero_grad(set_to_none=True)


def d_fwd_bwd():
    set_grad(D, True)
    x = img.clone().requires_grad_(True)
    out = D(x, c)
    out.sum().backward()
    set_grad(D, False)
    D.zero_grad(set_to_none=True)


def aug_fwd():
    with torch.no_grad():
        aug(img)


def aug_fwd_bwd():
    x = img.clone().requires_grad_(True)
    aug(x).sum().backward()


for p in list(G.parameters()) + list(D.parameters()):
    p.grad = torch.zeros_like(p)
optG, optD = tr.phases[0].opt, tr.phases[2].opt

measure('G.mapping fwd+bwd', g_map_fwd_bwd)
measure('G.synthesis fwd (no grad)', g_syn_fwd)
measure('G.synthesis fwd+bwd', g_syn_fwd_bwd)
measure('D fwd+bwd (incl. grad wrt image)', d_fwd_bwd)
measure('AugmentPipe fwd', aug_fwd)
measure('AugmentPipe fwd+bwd', aug_fwd_bwd)
for p in list(G.parameters()) + list(D.parameters()):
    p.grad = torch.zeros_like(p)
measure('Adam step G', lambda: optG.step())
measure('Adam step D', lambda: optD.step())
measure('train_step (all)', lambda: tr.train_step(img * 127.5 + 127.5, c), warm=1)
