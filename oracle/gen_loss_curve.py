"""ORACLE tooling: fixed-seed loss curves of the REAL reference (CPU `ref` path) for `tests/golden/loss_curve.npz`.

North star: "a fixed-seed 1-kimg run must track the reference's G/D losses within a stated tolerance".  The reference's own
training loop cannot be imported (outbound webhook at import), so the loop below restates
S3/training/training_loop_mi_multimodal.py:243-255, 308-357 around the reference's OWN networks and loss: phases
Gmain / Greg / Dmain / Dreg with lazy-regularisation-adjusted Adam, per-iteration latents, gradient nan_to_num, step.
Randomness comes from oracle/det_rng.py so that any device reproduces it.  Initial weights: tests/golden/model.npz.
Run in the build container: `python oracle/gen_loss_curve.py`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import gen_golden as gg  # noqa: E402
from oracle.det_rng import deterministic_rng  # noqa: E402

BATCH = 8
ITERS = 128                      # 1.024 kimg
SEED = 1234
LOSS_KW = dict(r1_gamma=0.4096, style_mixing_prob=0.9, pl_weight=2, pl_no_weight_grad=True)
OPT = dict(lr=0.0025, betas=[0, 0.99], eps=1e-8)
G_REG, D_REG = 4, 16
STATS = ['Loss/G/loss', 'Loss/D/loss', 'Loss/pl_penalty', 'Loss/r1_penalty', 'Loss/scores/fake', 'Loss/scores/real']


def run_curve(networks, loss_mod, training_stats, golden_model, device='cpu', probe=None, iters=ITERS, G_kw=None, D_kw=None, loss_extra=None):
    """One fixed-seed run; `networks` / `loss_mod` / `training_stats` are either the reference's modules or ours."""
    dev = torch.device(device)
    torch.manual_seed(SEED)            # (initial weights when no golden state is given; the CPU generator, so device-independent)
    G = networks.Generator(**(G_kw or gg.G_KW)).train().requires_grad_(False)
    D = networks.Discriminator(**(D_kw or gg.D_KW)).train().requires_grad_(False)
    if golden_model is not None:
        G.load_state_dict({k[len('model/G/'):]: torch.from_numpy(golden_model[k]) for k in golden_model.files if k.startswith('model/G/')})
        D.load_state_dict({k[len('model/D/'):]: torch.from_numpy(golden_model[k]) for k in golden_model.files if k.startswith('model/D/')})
    G, D = G.to(dev), D.to(dev)
    loss = loss_mod.StyleGAN2Loss(device=dev, G=G, D=D, augment_pipe=None, **LOSS_KW, **(loss_extra or {}))     # loss_extra: schedule options of OUR loss only
    phases = []
    for name, module, interval in [('G', G, G_REG), ('D', D, D_REG)]:
        ratio = interval / (interval + 1)
        opt = torch.optim.Adam(module.parameters(), lr=OPT['lr'] * ratio, betas=[b ** ratio for b in OPT['betas']], eps=OPT['eps'])
        phases += [(name + 'main', module, opt, 1), (name + 'reg', module, opt, interval)]
    log = {k: [] for k in STATS}
    orig_report = training_stats.report

    def report(name, value):
        if name in log:
            log[name].append(float(torch.as_tensor(value).detach().float().mean().cpu()))
        return value
    training_stats.report = report
    loss_mod.training_stats.report = report
    np.random.seed(SEED)
    try:
        with deterministic_rng(SEED):
            for it in range(iters):
                real = (torch.rand([BATCH, 1, 32, 32]) * 2 - 1).to(dev)
                real_c = torch.nn.functional.one_hot(torch.from_numpy(np.random.randint(2, size=BATCH)), 2).float().to(dev)
                zs = torch.randn([len(phases) * BATCH, G.z_dim]).to(dev).split(BATCH)
                cs = torch.nn.functional.one_hot(torch.from_numpy(np.random.randint(2, size=len(phases) * BATCH)), 2).float().to(dev).split(BATCH)
                for (name, module, opt, interval), z, c in zip(phases, zs, cs):
                    if it % interval != 0:
                        continue
                    opt.zero_grad(set_to_none=True)
                    module.requires_grad_(True)
                    loss.accumulate_gradients(phase=name, real_img=real, real_c=real_c, gen_z=z, gen_c=c, gain=interval, cur_nimg=it * BATCH)
                    module.requires_grad_(False)
                    if probe is not None:
                        probe(it, name, module)
                    for p in module.parameters():
                        # Second-order gradients that are structurally zero (biases under lrelu in the R1 / path-length
                        # passes) come back as ZERO tensors from the reference's `ref` path (plain autograd through
                        # leaky_relu) but as None from its CUDA plugin path (OPS/bias_act.py:196-204, has_2nd_grad False)
                        # -- and Adam treats the two differently (a zero gradient still decays the second moment).  The
                        # product mirrors the CUDA plugin; normalise both sides to None so the curves are comparable.
                        if p.grad is not None and float(p.grad.abs().max()) == 0.0:
                            p.grad = None
                        if p.grad is not None:
                            torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
                    opt.step()
    finally:
        training_stats.report = orig_report
        loss_mod.training_stats.report = orig_report
    return {k: np.asarray(v, dtype=np.float64) for k, v in log.items()}


def main():
    ref = gg.import_reference()
    from torch_utils import training_stats as ref_stats          # the reference's module (S3 is on sys.path)
    torch.set_num_threads(8)
    golden_model = np.load(os.path.join(gg.OUT, 'model.npz'))
    curve = run_curve(ref.networks, ref.loss, ref_stats, golden_model)
    out = {'curve/' + k: v for k, v in curve.items()}
    out['meta/batch'], out['meta/iters'], out['meta/seed'] = np.asarray(BATCH), np.asarray(ITERS), np.asarray(SEED)
    path = os.path.join(gg.OUT, 'loss_curve.npz')
    np.savez_compressed(path, **out)
    for k, v in curve.items():
        print(f'{k:20s} n={len(v):4d} first {v[:3]} last {v[-3:]}')
    print(f'wrote {path} ({os.path.getsize(path)} bytes)')


if __name__ == '__main__':
    main()
