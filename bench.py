#!/usr/bin/env python
"""bench.py -- train kimg/s of the StyleGAN2-ADA hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                 our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W  the reference's CPU (`ref`) path, oracle port

A "step" is one training iteration of Gan-track's loop (S3/training/training_loop_mi_multimodal.py:308-376): the
phases Gmain, Greg (every 4th), Dmain, Dreg (every 16th) with gradient exchange, Adam, G_ema and the ADA update,
over one global batch of synthetic CLARO-shaped slices (float32 [B,1,256,256] in [0,255], 2-class one-hot labels).
Workload at N=1: BASELINE.json configs[1] (256x256 1-ch, batch 32, cbase 16384, map-depth 8, fp16 top-4 resolutions,
lazy R1 + path-length).  N>1: the same per-GPU batch on every rank (weak scaling), one NCCL gradient all-reduce per phase.

One JSON line on stdout (rank 0).  `value`: inputs already resident in HBM.  `e2e`: through `Trainer.train_step` with
pinned HOST inputs copied H2D every step and a scalar statistic read back D2H every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--res', type=int, default=256)
    ap.add_argument('--batch-gpu', type=int, default=32, help='images per GPU per step (weak scaling: the default, global batch = 32 x N)')
    ap.add_argument('--global-batch', type=int, default=None, help='fixed GLOBAL batch split over the N ranks (strong scaling, BASELINE configs 3 / 4: '
                    'batch 64 over 2/4/8 GPUs, batch_gpu = B // N as S3/train_mi_multimodal.py:260); overrides --batch-gpu')
    ap.add_argument('--no-igemm', action='store_true', help='A/B switch: route every convolution to the library (conv_backend.allow_igemm = False)')
    ap.add_argument('--no-library', action='store_true', help='fail instead of routing any convolution to the library (conv_backend.allow_library = False)')
    ap.add_argument('--cbase', type=int, default=None)
    ap.add_argument('--aug', default='ada', choices=['ada', 'noaug'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--fuse-bias-act', action='store_true', help="A/B: the discriminator's bias_act runs in EVERY convolution epilogue (conv2d_gradfix.fuse_bias_act = True; default 'auto')")
    ap.add_argument('--no-fuse-residual', action='store_true', help="A/B: the discriminator's residual add as its own pass (conv2d_gradfix.fuse_residual = False)")
    ap.add_argument('--no-fuse-bias-act', action='store_true', help="A/B: bias_act always as its own pass (conv2d_gradfix.fuse_bias_act = False)")
    ap.add_argument('--no-rooflines', action='store_true', help='skip the per-kernel roofline microbenchmarks (scaling runs)')
    ap.add_argument('--no-overlap', action='store_true')
    ap.add_argument('--graph-overlap', action='store_true', help='capture the bucketed gradient exchange inside each phase graph (opt-in, see Trainer.graph_overlap)')
    ap.add_argument('--two-pass-dmain', action='store_true', help='score generated and real images in two discriminator passes (the reference schedule) instead of one merged pass')
    ap.add_argument('--no-graphs', action='store_true', help='launch every kernel from Python instead of replaying per-phase CUDA graphs')
    ap.add_argument('--cpu-batch', type=int, default=4)
    ap.add_argument('--profile-range', action='store_true', help='cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p['hbm_gbs'], tflops=p['bf16_tflops'], tflops_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']), src='measured')
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, src='fallback')


# ------------------------------------------------------------------------------------------------ clocks sampling

class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)

def cpu_iteration_runner(res, cbase, batch, aug):
    """Build the training step on the CPU with the oracle's primitive ops (= the reference's `impl='ref'` path) and
    return (run_phase, phases).  This is the only place bench.py executes oracle/."""
    from oracle.backend import oracle_ops
    from gan_track_b200.training import training_loop as tl
    cfg = tl.claro_config(resolution=res, batch=batch, num_gpus=1, cbase=cbase, aug=aug)
    ctx = oracle_ops()
    ctx.__enter__()
    trainer = tl.Trainer(cfg, rank=0, device='cpu', overlap=False)
    real = torch.rand([batch, 1, res, res]) * 255
    idx = torch.randint(0, 2, [batch])
    real_c = torch.nn.functional.one_hot(idx, 2).float()
    return trainer, real, real_c, ctx


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port; the Python reference itself
    cannot travel to the GPU box), all host threads, batch `--cpu-batch` per step, real phase schedule."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    res = args.res
    cbase = args.cbase or (16384 if res <= 256 else 32768)
    trainer, real, real_c, ctx = cpu_iteration_runner(res, cbase, args.cpu_batch, args.aug)
    try:
        for _ in range(args.warmup):
            trainer.train_step(real, real_c)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            trainer.train_step(real, real_c)
        dt = time.perf_counter() - t0
    finally:
        ctx.__exit__(None, None, None)
    kimg_s = args.cpu_batch * args.steps / 1000.0 / dt
    sample = f'{args.steps} training iterations at batch {args.cpu_batch}, {res}x{res} 1-ch, real phase schedule (after {args.warmup} warm-up)'
    line = {
        'impl': 'reference', 'metric': 'train kimg/s, StyleGAN2-ADA %dx%d 1-ch' % (res, res), 'value': kimg_s, 'unit': 'kimg/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1000.0, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'claro_stylegan2-ada shape {res}x{res} 1-ch cbase {cbase} map-depth 8, ADA={args.aug}; CPU port: the oracle restatement of the '
                               f"reference's ref-impl ops (oracle/ops_ref.py) under this repository's host code (the Python reference cannot travel to the GPU box)",
                   'global_batch': args.cpu_batch},
        'cpu_baseline': {'value': kimg_s, 'unit': 'kimg/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': kimg_s, 'unit': 'kimg/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(res, cbase, aug, batch=4):
    """Bounded sample for the `cpu_baseline` object of our arm: one warm-up iteration + iterations until ~12 s."""
    cores = os.cpu_count() or 1
    old = torch.get_num_threads()
    torch.set_num_threads(cores)
    trainer, real, real_c, ctx = cpu_iteration_runner(res, cbase, batch, aug)
    try:
        trainer.train_step(real, real_c)            # iteration 0 runs all four phases; used as warm-up
        # Amortised iteration = Gmain + Dmain every step, Greg every 4th, Dreg every 16th: time 16-step-equivalent by
        # timing each kind of iteration once.
        times = {}
        t0 = time.perf_counter(); trainer.train_step(real, real_c); times['main'] = time.perf_counter() - t0      # idx 1: Gmain+Dmain
        trainer.batch_idx = 4
        t0 = time.perf_counter(); trainer.train_step(real, real_c); times['main+Greg'] = time.perf_counter() - t0
        trainer.batch_idx = 16
        t0 = time.perf_counter(); trainer.train_step(real, real_c); times['all'] = time.perf_counter() - t0
    finally:
        ctx.__exit__(None, None, None)
        torch.set_num_threads(old)
    t16 = 12 * times['main'] + 3 * times['main+Greg'] + 1 * times['all']
    kimg_s = batch * 16 / 1000.0 / t16
    sample = (f'batch {batch}, {res}x{res}: one iteration of each kind timed (main {times["main"]:.2f}s, +Greg {times["main+Greg"]:.2f}s, '
              f'+Greg+Dreg {times["all"]:.2f}s), combined with the 16-iteration schedule 12/3/1')
    return {'value': kimg_s, 'unit': 'kimg/s', 'cores': cores, 'kind': 'port', 'sample': sample}


# ------------------------------------------------------------------------------------------------ our arm

# `roofline` reports the hot kernel with the largest share of the step, DERIVED from the run: every convolution call of one replay of
# each captured phase is logged by shape (conv_igemm.call_log), multiplied by the phase's replays in the timed region and by the
# per-launch time measured live in `hot_kernel_rooflines`; the case below is only the fallback when graphs are off.
DOMINANT_FALLBACK = 'conv_igemm_halo fwd 3x3 64->64 @256x256'


def dominant_kernel(trainer, phase_counts, roof_all, batch_gpu):
    """-> (case name, {case: ms per step}) over the convolution cases of `roof_all` (measured at batch 32: scaled by batch)."""
    per_step = {}
    steps = max(1, max(phase_counts.values()) if phase_counts else 1)
    for ph in trainer.phases:
        shapes = ph.get('conv_shapes')
        if not shapes:
            continue
        for (kind, n, cin, cout, h, w, k, stride, transpose), calls in shapes.items():
            if k != 3 or stride != 1 or transpose:
                continue
            name = (f'conv_igemm_halo fwd 3x3 {cin}->{cout} @{h}x{w}' if kind == 'fwd' else f'conv_wgrad_halo 3x3 {cin}x{cout} over 32x{h}x{w} pixels')
            if name in roof_all:
                per_step[name] = per_step.get(name, 0.0) + calls * phase_counts.get(ph.name, 0) / steps * roof_all[name]['ms_per_launch'] * n / 32.0
    if not per_step:
        return DOMINANT_FALLBACK, {}
    return max(per_step, key=per_step.get), {k: round(v, 3) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])}


def hot_kernel_rooflines(device, pk):
    """Rooflines of OUR hot kernels at their training shapes (batch 32 of the 256x256 config), timed live with CUDA
    events on the launching stream.  Cold caches without an interfering flush: each case rotates over enough distinct
    input/output buffer sets that consecutive launches never touch the same memory within 2x the 126 MB L2, and a batch
    of launches is timed back to back (sustained figure; the denominators are MEASURED_PEAKS.json's HBM copy bandwidth
    and bf16 burst throughput).  Algorithmic bytes / flops per launch: DESIGN.md section 3."""
    from gan_track_b200.torch_utils.ops import bias_act, conv_igemm, modulated, upfirdn2d
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f)
        except Exception:
            traffic = {}
    cl = torch.channels_last
    L2 = 126 << 20

    def rot(make, nbytes_per_set):
        n = max(2, int(np.ceil(2 * L2 / max(nbytes_per_set, 1))) + 1)
        return [make() for _ in range(min(n, 12))]

    def t16(shape):
        return torch.randn(shape, device=device).to(torch.float16).contiguous(memory_format=cl)

    def timeit(fns, iters=24):
        for f in fns:
            f()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(iters):
            fns[i % len(fns)]()
        e.record()
        e.synchronize()
        return s.elapsed_time(e) / iters

    out = {}

    def add(name, kernel, bound, work, ms):
        if bound == 'tensor':
            achieved, peak, unit = work / (ms * 1e-3) / 1e12, pk['tflops'], 'TFLOP/s'
            extra = {'algorithmic_flops_per_launch': work}
        else:
            achieved, peak, unit = work / (ms * 1e-3) / 1e9, pk['hbm_gbs'], 'GB/s'
            extra = {'algorithmic_bytes_per_launch': work}
        out[name] = dict(kernel=kernel, bound=bound, achieved=achieved, peak=peak, unit=unit, frac=achieved / peak, traffic=traffic.get(name),
                         peak_source=pk['src'], ms_per_launch=ms, **extra)

    cfg = dict(output_padding=(0, 0), groups=1, stride=(1, 1), padding=(1, 1))
    for ci, co, r, tag in [(512, 512, 32, 'conv_igemm_halo2_kernel<256,1,10> (CTA pair)'), (256, 256, 64, 'conv_igemm_halo2_kernel<256,1,10> (CTA pair)'),
                           (128, 128, 128, 'conv_igemm_halo2_kernel<128,2,12> (CTA pair)'), (64, 64, 256, 'conv_rows_kernel (row streaming, N = 192)')]:
        xs = rot(lambda: t16([32, ci, r, r]), 32 * ci * r * r * 2 * 2)
        w = (torch.randn([co, ci, 3, 3], device=device) / (ci * 9) ** 0.5).to(torch.float16)
        pkd = conv_igemm.pack_weight(w, False)
        fl = 2.0 * 32 * r * r * ci * co * 9
        ms = timeit([lambda x=x: conv_igemm.igemm_forward(x, w, transpose=False, packed=pkd, **cfg) for x in xs])
        add(f'conv_igemm_halo fwd 3x3 {ci}->{co} @{r}x{r}', tag, 'tensor', fl, ms)
        ms = timeit([lambda x=x: conv_igemm.igemm_wgrad(x, x, (co, ci, 3, 3), transpose=False, **cfg) for x in xs])
        add(f'conv_wgrad_halo 3x3 {ci}x{co} over 32x{r}x{r} pixels',
            'conv_wgrad_halo_wide_kernel<1> + wgrad_reduce_ab_kernel' if co % 128 == 0 else 'conv_wgrad_halo_kernel<1> + wgrad_reduce_kernel', 'tensor', fl, ms)
        del xs

    f = upfirdn2d.setup_filter([1, 3, 3, 1], device=device)
    xs = rot(lambda: t16([32, 64, 257, 257]), 32 * 64 * 257 * 257 * 4)
    ms = timeit([lambda x=x: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4) for x in xs])
    add('upfirdn2d 4x4 blur [32,64,257,257]->[32,64,256,256] f16 NHWC', 'upfirdn2d_tma_kernel<half>', 'hbm', (32 * 64 * 257 * 257 + 32 * 64 * 256 * 256) * 2, ms)
    del xs
    xs = rot(lambda: t16([32, 64, 256, 256]), 32 * 64 * 256 * 256 * 2 * 5 // 4)
    ms = timeit([lambda x=x: upfirdn2d.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1]) for x in xs])
    add('upfirdn2d 4x4 down=2 [32,64,256,256]->[32,64,128,128] f16 NHWC', 'upfirdn2d_tma_down2_kernel<half>', 'hbm', 32 * 64 * (256 * 256 + 128 * 128) * 2, ms)
    del xs
    xs = rot(lambda: t16([32, 64, 128, 128]), 32 * 64 * 128 * 128 * 2 * 5)
    ms = timeit([lambda x=x: upfirdn2d.upfirdn2d(x, f, up=2, padding=[2, 1, 2, 1], gain=4) for x in xs])
    add('upfirdn2d 4x4 up=2 [32,64,128,128]->[32,64,256,256] f16 NHWC', 'upfirdn2d_tma_up2_kernel<half,0,0>', 'hbm', 32 * 64 * (256 * 256 + 128 * 128) * 2, ms)
    del xs

    shape = [32, 64, 256, 256]
    nb = 32 * 64 * 256 * 256 * 2
    xs = rot(lambda: t16(shape), 2 * nb)
    b = torch.randn([64], device=device, dtype=torch.float16)
    ms = timeit([lambda x=x: bias_act.bias_act(x, b, act='lrelu', clamp=256.0) for x in xs])
    add('bias_act fwd lrelu [32,64,256,256] f16 NHWC', 'bias_act_bulk_kernel<half,lrelu>', 'hbm', 2 * nb, ms)
    lib = _lib_handle()
    ws = torch.empty([lib.gt_bias_act_bwd_workspace(32 * 256 * 256, 64, 1)], dtype=torch.float32, device=device)
    db = torch.empty([64], dtype=torch.float32, device=device)
    outs = [torch.empty_like(x) for x in xs]
    from gan_track_b200 import _lib

    def bwd(i):
        dy, y, dx = xs[i], xs[(i + 1) % len(xs)], outs[i]
        _lib.check(lib.gt_bias_act_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(dx), _lib.ptr(db), _lib.ptr(ws), ws.numel(), 1, 3, 0.2, float(np.sqrt(2)), 256.0,
                                       32 * 256 * 256, 64, 1, _lib.stream_of(dy)), 'gt_bias_act_bwd')
    ms = timeit([lambda i=i: bwd(i) for i in range(len(xs))])
    add('bias_act bwd (dx + db) [32,64,256,256] f16 NHWC', 'bias_act_bwd_bulk_kernel<half,lrelu> + reduce_partials_kernel', 'hbm', 3 * nb, ms)
    del outs
    sm = torch.randn([32, 64], device=device)
    dm = torch.rand([32, 64], device=device)
    nz = torch.randn([32, 1, 256, 256], device=device).to(torch.float16)
    ms = timeit([lambda x=x: modulated.mod_scale(x, sm) for x in xs])
    add('mod_scale fwd [32,64,256,256] f16 NHWC', 'mod_scale_fwd_bulk_kernel<half>', 'hbm', 2 * nb, ms)
    ms = timeit([lambda x=x: modulated.demod_act(x, dm, nz, b, act='lrelu', gain=float(np.sqrt(2)), clamp=256.0) for x in xs])
    add('demod_act fwd [32,64,256,256] f16 NHWC', 'demod_act_fwd_bulk_kernel<half,lrelu>', 'hbm', 2 * nb, ms)
    # backward passes through autograd on retained graphs (per-sample style / demodulation / noise / bias gradients included)
    graphs = []
    for x in xs[:3]:
        xr, sr = x.detach().requires_grad_(True), sm.clone().requires_grad_(True)
        graphs.append((modulated.mod_scale(xr, sr), [xr, sr]))
    dy = torch.randn_like(xs[0])
    ms = timeit([lambda g=g: torch.autograd.grad(g[0], g[1], dy, retain_graph=True) for g in graphs])
    add('mod_scale bwd (gx + gs) [32,64,256,256] f16 NHWC', 'mod_scale_bwd_bulk_kernel<half>', 'hbm', 3 * nb, ms)
    graphs = []
    for x in xs[:3]:
        xr, dr, nr, br = x.detach().requires_grad_(True), dm.clone().requires_grad_(True), nz.clone().requires_grad_(True), b.clone().requires_grad_(True)
        graphs.append((modulated.demod_act(xr, dr, nr, br, act='lrelu', gain=float(np.sqrt(2)), clamp=256.0), [xr, dr, nr, br]))
    ms = timeit([lambda g=g: torch.autograd.grad(g[0], g[1], dy, retain_graph=True) for g in graphs])
    add('demod_act bwd (gx + gd + gnoise + gb) [32,64,256,256] f16 NHWC', 'demod_act_bwd_bulk_kernel<half,lrelu>', 'hbm', 4 * nb, ms)
    del graphs, xs
    return out


def _lib_handle():
    from gan_track_b200 import _lib
    return _lib.load()


def run_ours(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=device)
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)'

    from gan_track_b200 import _lib
    from gan_track_b200.torch_utils import training_stats
    from gan_track_b200.torch_utils.ops import conv_backend
    from gan_track_b200.training import training_loop as tl
    _lib.load()
    if world > 1:
        training_stats.init_multiprocessing(rank=rank, sync_device=device)

    res = args.res
    cbase = args.cbase or (16384 if res <= 256 else 32768)
    batch_gpu = args.batch_gpu
    scaling = 'weak'
    if args.global_batch is not None:
        assert args.global_batch % world == 0, '--global-batch must divide over the ranks'
        batch_gpu, scaling = args.global_batch // world, 'strong'
    global_batch = batch_gpu * world
    conv_backend.allow_igemm = not args.no_igemm
    from gan_track_b200.torch_utils.ops import conv2d_gradfix as _cg
    if args.fuse_bias_act:
        _cg.fuse_bias_act = True
    if args.no_fuse_bias_act:
        _cg.fuse_bias_act = False
    if args.no_fuse_residual:
        _cg.fuse_residual = False
    conv_backend.allow_library = not args.no_library
    from gan_track_b200.torch_utils.ops import conv_igemm
    conv_igemm.call_log = {}
    gamma = 0.0002 * res ** 2 / global_batch if res != 256 or global_batch != 32 else 0.4096
    cfg = tl.claro_config(resolution=res, batch=global_batch, num_gpus=world, cbase=cbase, aug=args.aug, gamma=gamma)
    trainer = tl.Trainer(cfg, rank=rank, device=device, overlap=not args.no_overlap, use_graphs=not args.no_graphs, merge_d_passes=not args.two_pass_dmain,
                         graph_overlap=args.graph_overlap)

    g = torch.Generator().manual_seed(1234 + rank)
    host_img = (torch.rand([batch_gpu, 1, res, res], generator=g) * 255).pin_memory()
    host_c = torch.nn.functional.one_hot(torch.randint(0, 2, [batch_gpu], generator=g), 2).float().pin_memory()
    dev_img, dev_c = host_img.to(device), host_c.to(device)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # Graph mode captures each phase at its second occurrence; Dreg runs every 16th iteration, so 17 warm-up iterations put
    # every capture before the timed region.
    n_warm = max(args.warmup, 3) if args.no_graphs else max(args.warmup, 17)
    if n_warm != args.warmup and rank == 0:
        print(f'bench.py: --warmup {args.warmup} raised to {n_warm}: every phase graph (Dreg runs every 16th iteration) is captured at its second '
              f'occurrence and must be captured before the timed region; the JSON line reports warmup={n_warm}', file=sys.stderr, flush=True)
    for _ in range(n_warm):
        trainer.train_step(dev_img, dev_c)

    # ---- timed: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    counts0 = dict(trainer.phase_counts)
    launches0 = _lib.launches
    conv0 = dict(conv_backend.stats)
    barrier()
    if rank == 0:
        sampler.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_range:
        torch.cuda.profiler.start()
    trainer.time_phases = True
    s.record()
    for _ in range(args.steps):
        trainer.train_step(dev_img, dev_c)
    e.record()
    barrier()
    phase_ms = {k: {'replays': n, 'ms': round(t, 3)} for k, (n, t) in trainer.phase_times().items()}
    trainer.time_phases = False
    if args.profile_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(s.elapsed_time(e))
    launches = _lib.launches - launches0
    conv_stats = {k: conv_backend.stats[k] - conv0[k] for k in conv0}
    phase_counts = {k: trainer.phase_counts[k] - counts0[k] for k in counts0}
    value = global_batch * args.steps / 1000.0 / (ms_total * 1e-3)

    # ---- timed: end to end through the public API with host inputs ----
    e2e = None
    if not args.no_e2e:
        stat = torch.zeros([], device=device)
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        d2h = 0
        for _ in range(args.steps):
            trainer.train_step(host_img, host_c)               # H2D of the step's real batch + labels inside train_step
            stat = trainer.loss.pl_mean + trainer.augment_pipe.p if trainer.augment_pipe is not None else trainer.loss.pl_mean
            _ = float(stat.item())                             # D2H read of a step result
            d2h += 4
        e2.record()
        barrier()
        ms_e2e = max_over_ranks(s2.elapsed_time(e2))
        e2e = {'value': global_batch * args.steps / 1000.0 / (ms_e2e * 1e-3), 'unit': 'kimg/s',
               'h2d_bytes_per_step': int(host_img.numel() * 4 + host_c.numel() * 4) * world, 'd2h_bytes_per_step': 4 * world}

    nccl_ms = None
    if world > 1:
        trainer.check_consistency()
        # stand-alone cost of the per-phase exchange: one sum-all-reduce of each module's flat fp32 gradient (what the reference sends,
        # S3/training/training_loop_mi_multimodal.py:340-351), CUDA events, max over ranks
        nccl_ms = {}
        for name, module in [('G', trainer.G), ('D', trainer.D)]:
            buf = torch.zeros([sum(p.numel() for p in module.parameters())], device=device)
            for _ in range(3):
                torch.distributed.all_reduce(buf)
            barrier()
            s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s3.record()
            for _ in range(10):
                torch.distributed.all_reduce(buf)
            e3.record()
            barrier()
            t = max_over_ranks(s3.elapsed_time(e3) / 10)
            nccl_ms[name] = {'bytes': buf.numel() * 4, 'ms': round(t, 4), 'bus_GBps': round(2 * (world - 1) / world * buf.numel() * 4 / (t * 1e-3) / 1e9, 1)}
            del buf

    pk = peaks()
    roof_all = hot_kernel_rooflines(device, pk) if (rank == 0 and not args.no_rooflines) else None
    roof = None
    if roof_all:
        dom, dom_ms = dominant_kernel(trainer, phase_counts, roof_all, batch_gpu)
        roof = dict(roof_all[dom], case=dom, derived_from='conv call log x phase replays x live per-launch time (ms per step by case)', ms_per_step_by_case=dom_ms)
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_base = cpu_baseline_sample(res, cbase, args.aug, batch=args.cpu_batch)
        except Exception as ex:  # the CPU leg must never take the GPU number down with it
            cpu_base = {'value': None, 'unit': 'kimg/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'failed: {ex!r}'}

    if rank == 0:
        flops_per_img = {256: 399e9, 512: 1597e9}.get(res)
        line = {
            'metric': 'train kimg/s, StyleGAN2-ADA %dx%d 1-ch' % (res, res), 'value': value, 'unit': 'kimg/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': n_warm, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None,
            'dtype': 'f16', 'data': 'synthetic',
            'config': {'workload': f'claro_stylegan2-ada shape: {res}x{res} 1-ch, batch {batch_gpu}/GPU, cbase {cbase}, map-depth 8, fp16 top-4 resolutions, '
                                   f'lazy R1 (every 16) + path-length (every 4), ADA={args.aug}', 'global_batch': global_batch, 'batch_gpu': batch_gpu,
                       'parallelism': f'dp{world}', 'phase_counts_in_timed_region': phase_counts,
                       'l2_policy': 'per-step working set (GBs of activations) far exceeds the 126 MB L2; no explicit flush in the step loop',
                       'conv_routes': conv_stats, 'allow_igemm': conv_backend.allow_igemm, 'allow_library': conv_backend.allow_library, 'fuse_bias_act': _cg.fuse_bias_act, 'fuse_residual': _cg.fuse_residual, 'cuda_graphs': not args.no_graphs, 'exchange': ('bucketed all-reduce on a side stream inside each phase graph, overlapped with backward' if trainer.graph_overlap
                                                                       else 'one all-reduce of the flat gradient between two half-graphs') if world > 1 else 'none (1 GPU)',
                       'nccl_allreduce_alone': nccl_ms, 'dmain_one_pass': not args.two_pass_dmain, 'phase_ms': phase_ms},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roof, 'cpu_baseline': cpu_base, 'roofline_all': roof_all,
        }
        if flops_per_img:
            line['model_tflops'] = value * 1000 * flops_per_img / 1e12
            line['model_tensor_frac'] = line['model_tflops'] / world / pk['tflops_sustained']
        print(json.dumps(line), flush=True)
    if world > 1:
        # drop the captured graphs before the communicator goes away, then leave without waiting on NCCL teardown
        torch.cuda.synchronize()
        sys.stdout.flush()
        del trainer
        import gc
        gc.collect()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference_arm(a)
    else:
        run_ours(a)
