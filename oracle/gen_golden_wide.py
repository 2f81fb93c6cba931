"""ORACLE tooling: golden vectors of the REAL reference at widths the tensor-core kernels cover (tests/golden/model_wide.npz).

The small golden model (oracle/gen_golden.py, channel_max 16) never reaches the tcgen05 implicit-GEMM kernels, which need
channel counts that are multiples of 64.  This generator runs the reference's own Generator / Discriminator / StyleGAN2Loss /
AugmentPipe (CPU `ref` path, which forces fp32: S3/training/networks_stylegan2.py:419-420, 607-608) at resolution 64 with
64- and 128-channel layers -- on a CUDA device the blocks at resolution >= 8 of the same configuration run in fp16 -- and
records the network outputs and the parameter gradients of all four loss phases (S3/training/loss.py:64-139) with ADA on.

To keep the fixture small and machine-independent:
  * parameters are NOT stored: both this script and the tests fill the networks from one numpy RandomState in
    parameter-name order (`fill_parameters`);
  * randomness inside the phases (z, noise, style mixing, path-length probes, augmentation parameters) comes from the
    device-independent stream of oracle/det_rng.py, so the CUDA path reproduces the draws;
  * a gradient tensor with more than MAX_KEEP elements is stored as a fixed strided sample (`sample_index`) plus its L2 norm.
Run in the build container: `python oracle/gen_golden_wide.py`."""
import os
import re
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import gen_golden as gg  # noqa: E402
from oracle.det_rng import deterministic_rng  # noqa: E402

G_KW = dict(z_dim=64, c_dim=2, w_dim=64, img_resolution=64, img_channels=1, channel_base=4096, channel_max=128,
            mapping_kwargs=dict(num_layers=2), fused_modconv_default='inference_only')
D_KW = dict(c_dim=2, img_resolution=64, img_channels=1, channel_base=4096, channel_max=128, block_kwargs=dict(), mapping_kwargs=dict(),
            epilogue_kwargs=dict(mbstd_group_size=4))
LOSS_KW = dict(r1_gamma=0.4096, style_mixing_prob=0.9, pl_weight=2, pl_no_weight_grad=True)
BATCH = 4
PARAM_SEED = 77
DATA_SEED = 78
PHASE_SEED = 500
MAX_KEEP = 4096
PHASES = [('Gmain', 'G', 1), ('Greg', 'G', 4), ('Dmain', 'D', 1), ('Dreg', 'D', 16)]
ADA_P = 0.6
PL_MEAN = 0.21


def fill_parameters(module, rs):
    """Deterministic, non-degenerate parameter values from a numpy stream, in name order: weights and the
    constant input ~ N(0,1) (their initial distribution), biases and noise strengths ~ 0.1 N(0,1) (they initialise to
    zero / one, which would leave code paths unexercised); affine biases keep their `bias_init` of one plus a perturbation."""
    with torch.no_grad():
        for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):      # by NAME: independent of registration order
            v = rs.standard_normal(tuple(p.shape)).astype(np.float32)
            if re.search(r'mapping\.fc\d+\.', name):
                v = v / 0.01             # the mapping layers run with lr_multiplier 0.01: parameters are stored divided by it
            if name.endswith('noise_strength'):
                v = v * 0.1
            elif name.endswith('bias'):
                v = v * 0.1 + (1.0 if 'affine' in name else 0.0)
            p.copy_(torch.from_numpy(np.asarray(v)).reshape(p.shape))
        for name, b in sorted(module.named_buffers(), key=lambda kv: kv[0]):
            if name.endswith('noise_const'):
                b.copy_(torch.from_numpy(rs.standard_normal(tuple(b.shape)).astype(np.float32)))


def sample_index(numel):
    """Indices (into the flattened tensor) that the fixture keeps: everything up to MAX_KEEP elements, else an odd stride."""
    if numel <= MAX_KEEP:
        return np.arange(numel)
    step = (numel // MAX_KEEP) | 1
    return np.arange(0, numel, step)[:MAX_KEEP]


def make_inputs():
    rs = np.random.RandomState(DATA_SEED)
    z = torch.from_numpy(rs.standard_normal([BATCH, G_KW['z_dim']]).astype(np.float32))
    c = torch.nn.functional.one_hot(torch.tensor([0, 1, 1, 0]), 2).float()
    # smooth-ish slices in [-1, 1] (a sum of a low-frequency pattern and noise), like normalised CT slices
    yy, xx = np.meshgrid(np.linspace(-1, 1, 64), np.linspace(-1, 1, 64), indexing='ij')
    real = np.stack([0.6 * np.sin(3 * xx * (i + 1) + yy) * np.cos(2 * yy - i) + 0.3 * rs.standard_normal([64, 64]) for i in range(BATCH)])
    real = torch.from_numpy(np.clip(real, -1, 1).astype(np.float32))[:, None]
    return z, c, real


def build(networks, G_kw=None, D_kw=None):
    G = networks.Generator(**(G_kw or G_KW)).train().requires_grad_(False)
    D = networks.Discriminator(**(D_kw or D_KW)).train().requires_grad_(False)
    rs = np.random.RandomState(PARAM_SEED)
    fill_parameters(G, rs)
    fill_parameters(D, rs)
    return G, D


def run_phases(loss_mod, augment_cls, G, D, z, c, real, device='cpu', loss_extra=None, aug_ctor=None):
    """Network outputs and the four phases' parameter gradients on one fixed random stream.  Returns {key: tensor}."""
    dev = torch.device(device)
    out = {}
    aug = (aug_ctor or (lambda: augment_cls(run_dir=None, batch_size=BATCH, **gg.AUG_KW)))().train().requires_grad_(False).to(dev)
    aug.p.copy_(torch.as_tensor(ADA_P))
    loss = loss_mod.StyleGAN2Loss(device=dev, G=G, D=D, augment_pipe=aug, **LOSS_KW, **(loss_extra or {}))
    with deterministic_rng(PHASE_SEED):
        out['G_train_random'] = G(z, c)                 # fresh per-layer noise from the stream
        out['D_real'] = D(real, c)
        for phase, which, gain in PHASES:
            module = G if which == 'G' else D
            loss.pl_mean.copy_(torch.as_tensor(PL_MEAN))
            module.requires_grad_(True)
            for p in module.parameters():
                p.grad = None
            loss.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, gain=gain, cur_nimg=0)
            module.requires_grad_(False)
            for name, p in module.named_parameters():
                if p.grad is not None:
                    out[f'{phase}/{name}'] = p.grad.detach().clone()
            out[f'{phase}/pl_mean'] = loss.pl_mean.detach().clone()
    G.eval()
    out['G_eval_const'] = G(z, c, noise_mode='const')
    G.train()
    return out


def main():
    ref = gg.import_reference()
    torch.set_num_threads(8)
    G, D = build(ref.networks)
    z, c, real = make_inputs()
    res = run_phases(ref.loss, ref.augment.AugmentPipe, G, D, z, c, real)
    out = {}
    for k, v in res.items():
        v = v.detach().cpu().numpy().astype(np.float32)
        if '/' in k and not k.endswith('pl_mean'):
            flat = v.reshape(-1)
            out['wide/' + k + '/sample'] = flat[sample_index(flat.size)]
            out['wide/' + k + '/norm'] = np.asarray(np.sqrt((flat.astype(np.float64) ** 2).sum()))
            out['wide/' + k + '/absmax'] = np.asarray(np.abs(flat).max())
        else:
            out['wide/' + k] = v
    path = os.path.join(gg.OUT, 'model_wide.npz')
    np.savez_compressed(path, **out)
    ng = sum(1 for k in out if k.endswith('/sample'))
    print(f'model_wide.npz: {len(out)} arrays ({ng} gradient tensors), {os.path.getsize(path) / 1024:.0f} KiB')
    for k in ['wide/G_train_random', 'wide/D_real', 'wide/G_eval_const']:
        print(k, out[k].shape, float(np.abs(out[k]).max()))


if __name__ == '__main__':
    main()
