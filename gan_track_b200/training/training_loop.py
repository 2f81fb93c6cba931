"""One data-parallel training iteration of StyleGAN2-ADA: the hot loop of
S3/training/training_loop_mi_multimodal.py:308-376, packaged as `Trainer.train_step()`.

What is kept from the reference (so that a fixed seed tracks its losses): seeding `seed * num_gpus + rank` (:166-167),
TF32 off (:169-170), phase list with lazy-regularisation-adjusted Adam (:243-255), per-iteration latent draw
`randn([len(phases) * batch, z_dim])` (:319-320), phase order and intervals (:326-328), gradient exchange
sum -> /N -> nan_to_num(0, +-1e5) (:340-350), G_ema lerp with ramp-up (:358-366), ADA adjustment every
`ada_interval` iterations (:373-376).

What is out of scope and therefore absent: dataset / DataLoader, snapshots, metrics, logging, pickling
(SURVEY.md section 2, rows 10-20).  `train_step` takes the real batch as an argument.

Multi-GPU (one process per GPU, NCCL): replicas are kept bit-identical by a single sum-all-reduce of each phase's flat
fp32 gradient; with `overlap=True` (default) the flat buffer is cut into buckets in reverse parameter order and each
bucket's all-reduce is launched on a side stream from a post-accumulate-grad hook as soon as its last gradient is
ready, so the exchange overlaps the rest of backward (the reference blocks after backward, :343-345).

CUDA graphs (`use_graphs=True`): an iteration is ~11 000 kernel launches, which the Python host cannot issue as fast as a
B200 executes them.  Each phase (Gmain / Greg / Dmain / Dreg) is therefore captured once -- zero_grad, forward, backward
(incl. the double backwards), gradient flattening, nan_to_num, Adam -- and replayed with static input buffers.  This is
possible because nothing on the path reads device data on the host any more (fused ADA warp with device-resident
margins, select-based style mixing, capturable Adam).  With several ranks the phase is cut into two graphs around one
eager NCCL all-reduce of the flat gradient (0.3 ms for 99 MB over NVLink against tens of ms of compute, so the
bucketed overlap of the eager path is not needed there).  The first occurrence of every phase runs eagerly (warm-up:
cuDNN plan selection, constant caches), the second is captured.
"""
import copy

import os

import numpy as np
import torch

from .. import dnnlib
from ..torch_utils import misc, training_stats
from ..torch_utils.ops import conv2d_gradfix, grid_sample_gradfix
from . import augment as augment_mod
from . import flat_optim
from . import loss as loss_mod
from . import networks_stylegan2 as networks


def claro_config(resolution=256, batch=32, num_gpus=1, cbase=16384, cmax=512, map_depth=8, cond=True, gamma=0.4096,
                 aug='ada', target=0.6, glr=0.0025, dlr=0.0025, mbstd_group=4, fp32=False, seed=0, img_channels=1):
    """Config `c` that `train_mi_multimodal.py` builds for the CLARO launch script
    (REF/src/bash/claro-*.sh:18; S3/train_mi_multimodal.py:225-334), minus dataset / metrics / snapshot options."""
    c = dnnlib.EasyDict()
    c.common = dnnlib.EasyDict(c_dim=2 if cond else 0, img_resolution=resolution, img_channels=img_channels)
    c.G_kwargs = dnnlib.EasyDict(z_dim=512, w_dim=512, mapping_kwargs=dnnlib.EasyDict(num_layers=map_depth), channel_base=cbase,
                                 channel_max=cmax, fused_modconv_default='inference_only')
    c.D_kwargs = dnnlib.EasyDict(block_kwargs=dnnlib.EasyDict(freeze_layers=0), mapping_kwargs=dnnlib.EasyDict(),
                                 epilogue_kwargs=dnnlib.EasyDict(mbstd_group_size=mbstd_group), channel_base=cbase, channel_max=cmax)
    c.G_opt_kwargs = dnnlib.EasyDict(lr=glr, betas=[0, 0.99], eps=1e-8)
    c.D_opt_kwargs = dnnlib.EasyDict(lr=dlr, betas=[0, 0.99], eps=1e-8)
    c.loss_kwargs = dnnlib.EasyDict(r1_gamma=gamma, style_mixing_prob=0.9, pl_weight=2, pl_no_weight_grad=True)
    c.G_reg_interval = 4
    c.D_reg_interval = 16
    c.num_gpus = num_gpus
    c.batch_size = batch
    c.batch_gpu = batch // num_gpus
    c.ema_kimg = batch * 10 / 32
    c.ema_rampup = 0.05
    c.random_seed = seed
    c.augment_kwargs = None
    c.augment_p = 0
    c.ada_target = None
    c.ada_interval = 4
    c.ada_kimg = 500
    if aug != 'noaug':
        c.augment_kwargs = dnnlib.EasyDict(xflip=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, xint_max=0.05, rotate_max=3 / 360,
                                           xfrac_std=0.05, scale_std=0.05, aniso_std=0.05)
        if aug == 'ada':
            c.ada_target = target
        else:
            c.augment_p = float(aug) if not isinstance(aug, str) else 0.2
    if fp32:
        c.G_kwargs.num_fp16_res = c.D_kwargs.num_fp16_res = 0
        c.G_kwargs.conv_clamp = c.D_kwargs.conv_clamp = None
    return c


class GradBucketReducer:
    """Sum-all-reduce of one module's gradients, bucketed in reverse parameter order and overlapped with backward."""

    def __init__(self, params, world_size, bucket_bytes=32 << 20, overlap=True):
        self.params = list(params)
        self.world_size = world_size
        self.overlap = overlap and world_size > 1
        self.on_cuda = len(self.params) > 0 and self.params[0].is_cuda
        self.comm_stream = torch.cuda.Stream() if (self.overlap and self.on_cuda) else None
        self.buckets = []          # lists of param indices, reverse order (last layers' grads are ready first)
        cur, cur_bytes = [], 0
        for idx in reversed(range(len(self.params))):
            cur.append(idx)
            cur_bytes += self.params[idx].numel() * 4
            if cur_bytes >= bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {}
        for b, idxs in enumerate(self.buckets):
            for i in idxs:
                self.bucket_of[i] = b
        self._pending = None
        self._handles = []
        self._hooks = []
        self.fire_counts = {}
        if self.overlap:
            for i, p in enumerate(self.params):
                was = p.requires_grad
                p.requires_grad_(True)           # hooks can only be attached while the tensor requires grad
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
                p.requires_grad_(was)
        self.active = False

    def _make_hook(self, i):
        def hook(param):
            self.fire_counts[i] = self.fire_counts.get(i, 0) + 1
            if not self.active:
                return
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def begin(self, expected):
        """Arm for one phase.  `expected`: {param index: number of times its gradient is accumulated in this phase}
        (a phase may run several backward passes, e.g. Dmain = fake + real), learned from `fire_counts` the first time
        the phase runs; None disables overlap for this pass."""
        self._handles = []
        self.fire_counts = {}
        if not self.overlap or expected is None:
            self.active = False
            return
        self._pending = [0] * len(self.buckets)
        for i, n in expected.items():
            self._pending[self.bucket_of[i]] += n
        self.active = True

    def _launch(self, b):
        grads = [self.params[i].grad for i in self.buckets[b] if self.params[i].grad is not None]
        if not grads:
            return
        flat = torch.cat([g.flatten() for g in grads])
        if self.comm_stream is not None:
            ready = torch.cuda.Event()
            ready.record()
            self.comm_stream.wait_event(ready)
            with torch.cuda.stream(self.comm_stream):
                torch.distributed.all_reduce(flat)
                flat.div_(self.world_size)
                torch.nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5, out=flat)
            if not torch.cuda.is_current_stream_capturing():
                flat.record_stream(self.comm_stream)      # (inside a capture `_handles` keeps the bucket alive until finish() has joined the side stream)
        else:                               # CPU / gloo (tests): same bucketing, synchronous exchange
            torch.distributed.all_reduce(flat)
            flat.div_(self.world_size)
            torch.nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5, out=flat)
        self._handles.append((b, flat, grads))

    def finish(self):
        """Reduce whatever was not launched from hooks, wait for the side stream, scatter results back into .grad."""
        self.active = False
        params = [p for p in self.params if p.grad is not None]
        if not params:
            return
        if self.overlap and self._handles:
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            done = set()
            for b, flat, grads in self._handles:
                for g, piece in zip(grads, flat.split([g.numel() for g in grads])):
                    g.copy_(piece.reshape(g.shape))
                done.update(self.buckets[b])
            params = [p for i, p in enumerate(self.params) if p.grad is not None and i not in done]
            if not params:
                return
        flat = torch.cat([p.grad.flatten() for p in params])
        if self.world_size > 1:
            torch.distributed.all_reduce(flat)
            flat /= self.world_size
        misc.nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5, out=flat)
        for p, g in zip(params, flat.split([p.numel() for p in params])):
            p.grad = g.reshape(p.shape)


class Trainer:
    def __init__(self, cfg, rank=0, device=None, overlap=True, ops_sanity=True, use_graphs=False, merge_d_passes=None, graph_overlap=False):
        self.cfg = cfg
        self.rank = rank
        self.num_gpus = cfg.num_gpus
        self.device = torch.device(device) if device is not None else torch.device('cuda', rank)
        np.random.seed(cfg.random_seed * self.num_gpus + rank)
        torch.manual_seed(cfg.random_seed * self.num_gpus + rank)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cudnn.benchmark = True
        conv2d_gradfix.enabled = True
        grid_sample_gradfix.enabled = True

        self.use_graphs = bool(use_graphs) and self.device.type == 'cuda'
        # Graph mode on several GPUs.  graph_overlap=True: the whole phase is ONE graph in which the bucketed gradient exchange runs on a
        # side stream, forked from the backward pass by the parameters' post-accumulate hooks, so that NCCL overlaps the rest of backward.
        # False (default): two graphs around one eager all-reduce of the flat gradient (the reference also exchanges once, after
        # backward: training_loop_mi_multimodal.py:340-351).  Measured on 2 x B200 (profiles/r02_scaling_notes.md): the exchange is
        # 0.22 ms of a 17 ms phase, the overlapped form is not faster (43.8 vs 42.5 ms/step: three extra passes over the gradient) and
        # processes with captured NCCL work did not shut down cleanly, so it stays opt-in.
        self.graph_overlap = bool(graph_overlap) and bool(overlap) and cfg.num_gpus > 1 and self.use_graphs
        if self.use_graphs and cfg.loss_kwargs.get('blur_fade_kimg', 0) != 0:
            # a captured phase freezes host-side scalars derived from cur_nimg (blur_sigma and the blur taps, loss.py) at their
            # capture-time value, whereas the reference fades them (S3/training/loss.py:70); the StyleGAN2 configs use 0
            raise ValueError('use_graphs=True needs blur_fade_kimg == 0 (the blur schedule is a host-side function of cur_nimg)')
        dev = self.device
        self.G = networks.Generator(**cfg.G_kwargs, **cfg.common).train().requires_grad_(False).to(dev)
        self.D = networks.Discriminator(**cfg.D_kwargs, **cfg.common).train().requires_grad_(False).to(dev)
        self.G_ema = copy.deepcopy(self.G).eval()
        self.augment_pipe = None
        self.ada_stats = None
        if cfg.augment_kwargs is not None and (cfg.augment_p > 0 or cfg.ada_target is not None):
            self.augment_pipe = augment_mod.AugmentPipe(**cfg.augment_kwargs).train().requires_grad_(False).to(dev)
            self.augment_pipe.p.copy_(torch.as_tensor(cfg.augment_p))
            if cfg.ada_target is not None:
                self.ada_stats = training_stats.Collector(regex='Loss/signs/real')
        if self.num_gpus > 1:
            for module in [self.G, self.D, self.G_ema, self.augment_pipe]:
                if module is not None:
                    for t in misc.params_and_buffers(module):
                        torch.distributed.broadcast(t, src=0)
        self.loss = loss_mod.StyleGAN2Loss(device=dev, G=self.G, D=self.D, augment_pipe=self.augment_pipe, **cfg.loss_kwargs)
        # Dmain as one discriminator pass over [generated, real] (loss.StyleGAN2Loss._d_main_merged): on by default on the GPU
        self.loss.merge_d_passes = (self.device.type == 'cuda') if merge_d_passes is None else bool(merge_d_passes)
        self.loss.merge_mapping_passes = self.loss.merge_d_passes

        # Graph mode: parameters live in flat buffers; gradient scrub + Adam and the G_ema lerp are single kernels (csrc/optim.cu)
        self.flat = None
        if self.use_graphs:
            self.flat = dnnlib.EasyDict(G=flat_optim.FlatParams(self.G), D=flat_optim.FlatParams(self.D), G_ema=flat_optim.FlatParams(self.G_ema))
        self.phases = []
        for name, module, opt_kwargs, reg_interval in [('G', self.G, cfg.G_opt_kwargs, cfg.G_reg_interval), ('D', self.D, cfg.D_opt_kwargs, cfg.D_reg_interval)]:
            params = list(module.parameters())
            reducer = GradBucketReducer(params, self.num_gpus, overlap=overlap)
            adam_extra = dict(capturable=True, foreach=True) if self.use_graphs else {}
            if reg_interval is None:
                opt = torch.optim.Adam(params, **opt_kwargs, **adam_extra)
                self.phases.append(dnnlib.EasyDict(name=name + 'both', module=module, opt=opt, interval=1, reducer=reducer, grad_set=None))
                if self.flat is not None:
                    self.phases[-1].flat_opt = flat_optim.FlatAdam(self.flat[name], lr=opt_kwargs['lr'], betas=opt_kwargs['betas'], eps=opt_kwargs['eps'])
            else:
                ratio = reg_interval / (reg_interval + 1)
                kw = dict(opt_kwargs)
                kw['lr'] = kw['lr'] * ratio
                kw['betas'] = [beta ** ratio for beta in kw['betas']]
                opt = torch.optim.Adam(params, **kw, **adam_extra)
                self.phases.append(dnnlib.EasyDict(name=name + 'main', module=module, opt=opt, interval=1, reducer=reducer, grad_set=None))
                self.phases.append(dnnlib.EasyDict(name=name + 'reg', module=module, opt=opt, interval=reg_interval, reducer=reducer, grad_set=None))
                if self.flat is not None:
                    fopt = flat_optim.FlatAdam(self.flat[name], lr=kw['lr'], betas=kw['betas'], eps=kw['eps'])
                    self.phases[-1].flat_opt = self.phases[-2].flat_opt = fopt
        self.cur_nimg = 0
        self.batch_idx = 0
        self.time_phases = False          # bench.py: CUDA-event time of every graphed phase replay (read with phase_times())
        self._phase_events = []
        self.phase_counts = {p.name: 0 for p in self.phases}
        for ph in self.phases:
            ph.update(graphs=None, static=None, eager_runs=0, replay_counts=None)

    # ------------------------------------------------------------------------------------------------------------
    def draw_labels(self, n):
        c_dim = self.cfg.common.c_dim
        if c_dim == 0:
            return torch.zeros([n, 0], device=self.device)
        idx = np.random.randint(c_dim, size=n)
        onehot = np.zeros([n, c_dim], dtype=np.float32)
        onehot[np.arange(n), idx] = 1
        t = torch.from_numpy(onehot)
        if self.device.type == 'cuda':
            t = t.pin_memory()
        return t.to(self.device, non_blocking=True)

    def train_step(self, real_img, real_c, normalized=False):
        """One iteration over this rank's share of the global batch.
        real_img: [batch_size / num_gpus, C, H, W]; real_c: one-hot labels [.., c_dim].
        normalized=False (the reference loop's contract): float32 in [0, 255] as the loader delivers it
        (S3/training/dataset_mi_multimodal.py:259-264), mapped to [-1, 1] here like training_loop_mi_multimodal.py:317.
        normalized=True: already in [-1, 1] -- what `dataset.DeviceBatcher.batch()` / `.iterate()` return (the gather kernel
        applies the `/127.5 - 1` itself), so feed those with `normalized=True` or the reals would be normalised twice."""
        cfg, dev = self.cfg, self.device
        batch_gpu = cfg.batch_gpu
        real_img = real_img.to(dev, non_blocking=True).to(torch.float32)
        if not normalized:
            real_img = real_img / 127.5 - 1
        real_img = real_img.split(batch_gpu)
        real_c = real_c.to(dev, non_blocking=True).split(batch_gpu)
        # Like the reference (:319-323) every rank draws len(phases) * GLOBAL batch latents and uses the leading
        # batch_size / num_gpus of each phase's chunk (zip() below stops at the number of real micro-batches).
        all_gen_z = torch.randn([len(self.phases) * cfg.batch_size, self.G.z_dim], device=dev)
        all_gen_z = [z.split(batch_gpu) for z in all_gen_z.split(cfg.batch_size)]
        all_gen_c = self.draw_labels(len(self.phases) * cfg.batch_size)
        all_gen_c = [c.split(batch_gpu) for c in all_gen_c.split(cfg.batch_size)]

        for phase, phase_gen_z, phase_gen_c in zip(self.phases, all_gen_z, all_gen_c):
            if self.batch_idx % phase.interval != 0:
                continue
            self.phase_counts[phase.name] += 1
            if self.use_graphs:
                # one optimiser state per module: the graphed path steps `flat_opt`, the eager path torch's Adam -- never mix them
                if len(real_img) != 1:
                    raise RuntimeError(f'graph mode runs one micro-batch per phase: got {sum(len(r) for r in real_img)} images for batch_gpu={batch_gpu}')
                self._run_phase_graphed(phase, real_img[0], real_c[0], phase_gen_z[0], phase_gen_c[0])
            else:
                self._run_phase_eager(phase, real_img, real_c, phase_gen_z, phase_gen_c)

        # G_ema <- lerp(G, G_ema, beta)
        ema_nimg = cfg.ema_kimg * 1000
        if cfg.ema_rampup is not None:
            ema_nimg = min(ema_nimg, self.cur_nimg * cfg.ema_rampup)
        ema_beta = 0.5 ** (cfg.batch_size / max(ema_nimg, 1e-8))
        with torch.no_grad():
            if self.flat is not None:
                flat_optim.ema_update(self.flat.G_ema, self.flat.G, 1.0 - ema_beta)    # one launch over the flat buffers
            else:
                g_params = list(self.G.parameters())
                e_params = list(self.G_ema.parameters())
                torch._foreach_lerp_(e_params, g_params, 1.0 - ema_beta)            # p_ema = p.lerp(p_ema, beta)
            for b_ema, b in zip(self.G_ema.buffers(), self.G.buffers()):
                b_ema.copy_(b)

        self.cur_nimg += cfg.batch_size
        self.batch_idx += 1

        if self.ada_stats is not None and self.batch_idx % cfg.ada_interval == 0:
            self.ada_stats.update()
            adjust = np.sign(self.ada_stats['Loss/signs/real'] - cfg.ada_target) * (cfg.batch_size * cfg.ada_interval) / (cfg.ada_kimg * 1000)
            self.augment_pipe.p.copy_((self.augment_pipe.p + adjust).max(misc.constant(0, device=dev)))

    # ------------------------------------------------------------------------------------------------------------
    def _run_phase_eager(self, phase, real_img, real_c, phase_gen_z, phase_gen_c):
        phase.opt.zero_grad(set_to_none=True)
        phase.module.requires_grad_(True)
        single_round = len(real_img) == 1
        phase.reducer.begin(phase.grad_set if single_round else None)
        for r_img, r_c, g_z, g_c in zip(real_img, real_c, phase_gen_z, phase_gen_c):
            self.loss.accumulate_gradients(phase=phase.name, real_img=r_img, real_c=r_c, gen_z=g_z, gen_c=g_c, gain=phase.interval,
                                           cur_nimg=self.cur_nimg)
        phase.module.requires_grad_(False)
        if phase.grad_set is None and single_round and phase.reducer.fire_counts:
            phase.grad_set = dict(phase.reducer.fire_counts)
        phase.reducer.finish()
        phase.opt.step()

    # The two halves of a phase as they are captured.  Half A ends with the module's gradients flattened into one static
    # buffer (what the reference builds with torch.cat at :341-342); half B consumes it.
    def _phase_half_a(self, phase, st):
        phase.opt.zero_grad(set_to_none=True)
        phase.module.requires_grad_(True)
        if self.graph_overlap:
            phase.reducer.begin(phase.grad_set)          # None on the first (eager) occurrence: the hooks only count
        self.loss.accumulate_gradients(phase=phase.name, real_img=st.real_img, real_c=st.real_c, gen_z=st.gen_z, gen_c=st.gen_c,
                                       gain=phase.interval, cur_nimg=self.cur_nimg)
        phase.module.requires_grad_(False)
        if self.graph_overlap:
            if phase.grad_set is None and phase.reducer.fire_counts:
                phase.grad_set = dict(phase.reducer.fire_counts)
            phase.reducer.finish()                       # joins the side stream; .grad now holds sum / num_gpus, scrubbed
        st.with_grad = [p for p in phase.module.parameters() if p.grad is not None]
        st.active_idx = [i for i, p in enumerate(phase.module.parameters()) if p.grad is not None]
        st.flat = torch.cat([p.grad.flatten() for p in st.with_grad])

    def _phase_half_b(self, phase, st):
        if phase.get('flat_opt') is not None:
            # /num_gpus, nan_to_num and Adam in one kernel over the flat gradient (parameters without a gradient are skipped)
            phase.flat_opt.step(phase.name, st.active_idx, st.flat, grad_scale=1.0 if self.graph_overlap else 1.0 / self.num_gpus)
            return
        flat = st.flat
        if self.num_gpus > 1 and not self.graph_overlap:
            flat = flat / self.num_gpus
        flat = misc.nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5)
        for p, g in zip(st.with_grad, flat.split([p.numel() for p in st.with_grad])):
            p.grad = g.reshape(p.shape)
        phase.opt.step()

    def _run_phase_graphed(self, phase, r_img, r_c, g_z, g_c):
        from .. import _lib
        from ..torch_utils.ops import conv_backend
        if phase.static is None:
            phase.static = dnnlib.EasyDict(real_img=torch.empty_like(r_img), real_c=torch.empty_like(r_c), gen_z=torch.empty_like(g_z),
                                           gen_c=torch.empty_like(g_c), with_grad=None, flat=None)
        st = phase.static
        st.real_img.copy_(r_img)
        st.real_c.copy_(r_c)
        st.gen_z.copy_(g_z)
        st.gen_c.copy_(g_c)
        if phase.graphs is None and phase.eager_runs < 1:         # warm-up occurrence: eager
            self._phase_half_a(phase, st)
            if self.num_gpus > 1 and not self.graph_overlap:
                torch.distributed.all_reduce(st.flat)
            self._phase_half_b(phase, st)
            phase.eager_runs += 1
            return
        if phase.graphs is None:                                  # capture (does not execute), then replay below
            launches0, conv0 = _lib.launches, dict(conv_backend.stats)
            from ..torch_utils.ops import conv_igemm
            log0 = dict(conv_igemm.call_log) if conv_igemm.call_log is not None else None
            torch.cuda.synchronize()
            ga = torch.cuda.CUDAGraph()
            if self.num_gpus == 1 or self.graph_overlap:
                with torch.cuda.graph(ga):
                    self._phase_half_a(phase, st)
                    self._phase_half_b(phase, st)
                phase.graphs = (ga, None)
            else:
                with torch.cuda.graph(ga):
                    self._phase_half_a(phase, st)
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb, pool=ga.pool()):
                    self._phase_half_b(phase, st)
                phase.graphs = (ga, gb)
            # the Python-side launch / route counters only tick at capture time; remember the per-replay amounts
            phase.replay_counts = (_lib.launches - launches0, {k: conv_backend.stats[k] - conv0[k] for k in conv0})
            if log0 is not None:        # convolution shapes of one replay of this phase (bench.py derives the dominant kernel from it)
                phase.conv_shapes = {k: v - log0.get(k, 0) for k, v in conv_igemm.call_log.items() if v != log0.get(k, 0)}
                conv_igemm.call_log.clear()
                conv_igemm.call_log.update(log0)
            _lib.launches = launches0
            conv_backend.stats.update(conv0)
        ga, gb = phase.graphs
        ev = None
        if self.time_phases:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        ga.replay()
        if gb is not None:
            torch.distributed.all_reduce(st.flat)
            gb.replay()
        if ev is not None:
            ev[1].record()
            self._phase_events.append((phase.name, ev))
        n, routes = phase.replay_counts
        _lib.launches += n
        for k, v in routes.items():
            conv_backend.stats[k] += v

    def phase_times(self):
        """{phase: (replays, mean ms)} of the replays recorded since `time_phases` was switched on (synchronises)."""
        torch.cuda.synchronize(self.device)
        acc = {}
        for name, (a, b) in self._phase_events:
            n, t = acc.get(name, (0, 0.0))
            acc[name] = (n + 1, t + a.elapsed_time(b))
        self._phase_events = []
        return {k: (n, t / n) for k, (n, t) in acc.items()}

    def check_consistency(self):
        for module in (self.G, self.D):
            misc.check_ddp_consistency(module, ignore_regex=r'.*\.[^.]+_(avg|ema)')
