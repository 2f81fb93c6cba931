"""ORACLE tooling: generate the golden vectors under tests/golden/ by running the REAL reference on the CPU.

Run in the build container only (`python oracle/gen_golden.py`); `/root/reference` does not exist on the GPU box, so
tests read the committed .npz files, never the reference.  The reference is imported read-only from
/root/reference/src/models/stylegan3 with two compatibility shims that live here, not in the reference:
  1. empty `matplotlib`, `matplotlib.pyplot`, `openpyxl` modules so `training.augment_mi` imports;
  2. `torch._C._jit_get_operation` unwrapped from the (op, overloads) tuple torch 2.x returns, so that
     `grid_sample_gradfix` double-backward works (OPS/grid_sample_gradfix.py:60-65).
It never imports `train_mi_multimodal` / `training_loop_mi_multimodal` (outbound webhook).
"""
import os
import sys
import types

import numpy as np
import torch

S3 = '/root/reference/src/models/stylegan3'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def import_reference():
    sys.dont_write_bytecode = True
    if S3 not in sys.path:
        sys.path.insert(0, S3)
    for m in ['matplotlib', 'matplotlib.pyplot', 'openpyxl']:
        sys.modules.setdefault(m, types.ModuleType(m))
    orig = torch._C._jit_get_operation
    if not getattr(orig, '_gt_shim', False):
        class _CallableTuple(tuple):
            """(op, overload_names) as torch 2.x returns it -- torch's own callers unpack it -- that can also be CALLED like the
            bare op torch 1.x returned, which is what the reference does (OPS/grid_sample_gradfix.py:60-65)."""

            def __call__(self, *a, **k):
                return self[0](*a, **k)

        def shim(name):
            r = orig(name)
            return _CallableTuple(r) if isinstance(r, tuple) else r
        shim._gt_shim = True
        torch._C._jit_get_operation = shim
    import warnings
    warnings.filterwarnings('ignore')
    from torch_utils.ops import bias_act, conv2d_gradfix, conv2d_resample, fma, grid_sample_gradfix, upfirdn2d
    from training import augment_mi, loss, networks_stylegan2
    grid_sample_gradfix.enabled = True
    return types.SimpleNamespace(bias_act=bias_act, upfirdn2d=upfirdn2d, conv2d_resample=conv2d_resample, fma=fma,
                                 conv2d_gradfix=conv2d_gradfix, grid_sample_gradfix=grid_sample_gradfix,
                                 networks=networks_stylegan2, loss=loss, augment=augment_mi)


def npy(t):
    return t.detach().cpu().numpy()


# Shared case tables (imported by the tests so that generator and checker cannot drift apart) ----------------------

BIAS_ACT_PATH_CASES = [          # (name, act, gain, clamp, has_bias, shape)  -- the parameterisations on the path, SURVEY.md A.2
    ('lrelu_sqrt2_c256', 'lrelu', None, 256.0, True, (2, 8, 6, 6)),
    ('lrelu_g1_c181', 'lrelu', 1.0, 181.02, True, (2, 8, 6, 6)),
    ('linear_c256', 'linear', None, 256.0, True, (2, 1, 8, 8)),
    ('linear_sqrthalf_nobias', 'linear', float(np.sqrt(0.5)), None, False, (2, 8, 6, 6)),
    ('lrelu_fc', 'lrelu', None, None, True, (4, 16)),
    ('lrelu_sat', 'lrelu', 40.0, 3.0, True, (2, 4, 5, 5)),       # forces clamping so the masked-gradient branch is hit
]
ALL_ACTS = ['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish']

F4 = [1, 3, 3, 1]
SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466, 0.787641141030194,
        0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148]
UPFIRDN_CASES = [                # (name, filter, kwargs, x shape)  -- every (up, down, pad, flip, gain) on the path
    ('blur_up_pad1_g4', F4, dict(up=1, down=1, padding=[1, 1, 1, 1], gain=4), (2, 3, 9, 9)),
    ('blur_down_pad2', F4, dict(up=1, down=1, padding=[2, 2, 2, 2]), (2, 3, 8, 8)),
    ('img_up2', F4, dict(up=2, down=1, padding=[2, 1, 2, 1], gain=4), (2, 1, 8, 8)),
    ('skip_down2', F4, dict(up=1, down=2, padding=[1, 1, 1, 1]), (2, 3, 8, 8)),
    ('aug_up2_sym6', SYM6, dict(up=2, down=1, padding=[6, 5, 6, 5], gain=4), (2, 1, 14, 12)),
    ('aug_down2_sym6_flip', SYM6, dict(up=1, down=2, padding=[-1, -1, -1, -1], flip_filter=True), (2, 1, 28, 24)),
    ('odd_up3_down2', [1, 2, 4, 2, 1], dict(up=3, down=2, padding=[2, 0, 1, 3], gain=2), (1, 2, 7, 5)),
    ('crop_neg_pad', F4, dict(up=1, down=1, padding=[-1, 2, 0, -2]), (1, 2, 9, 10)),
    ('identity_none', None, dict(up=1, down=1, padding=0), (1, 2, 4, 4)),
]
CONV_RESAMPLE_CASES = [          # (name, Cin, Cout, k, kwargs, H)  -- the five cases of SURVEY.md A.4 plus generic fallbacks
    ('conv3', 4, 6, 3, dict(padding=1), 8),
    ('conv1', 4, 6, 1, dict(), 8),
    ('up2_k3', 4, 6, 3, dict(up=2, padding=1, flip_weight=False), 8),
    ('down2_k3', 4, 6, 3, dict(down=2, padding=1), 8),
    ('down2_k1', 4, 6, 1, dict(down=2), 8),
    ('up2_k1', 4, 6, 1, dict(up=2), 8),
    ('updown', 4, 6, 3, dict(up=2, down=2, padding=1), 8),
    ('asym_pad', 4, 6, 3, dict(padding=[1, 0, 2, 1]), 8),
]

G_KW = dict(z_dim=32, c_dim=2, w_dim=32, img_resolution=32, img_channels=1, channel_base=256, channel_max=16,
            mapping_kwargs=dict(num_layers=2), fused_modconv_default='inference_only')
D_KW = dict(c_dim=2, img_resolution=32, img_channels=1, channel_base=256, channel_max=16, block_kwargs=dict(), mapping_kwargs=dict(),
            epilogue_kwargs=dict(mbstd_group_size=4))
AUG_KW = dict(xflip=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, xint_max=0.05, rotate_max=3 / 360, scale_std=0.05, aniso_std=0.05,
              xfrac_std=0.05)                                          # REF/src/bash/claro-*.sh:18 + train_mi_multimodal.py:311-316
AUG_KW_FULL = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1,
                   hue=1, saturation=1, noise=1, cutout=1)
LOSS_KW = dict(r1_gamma=0.4096, style_mixing_prob=0.9, pl_weight=2, pl_no_weight_grad=True)


def gen_bias_act(ref, out):
    g = torch.Generator().manual_seed(1)
    for name, act, gain, clamp, has_b, shape in BIAS_ACT_PATH_CASES:
        x = (torch.randn(shape, generator=g) * 2).requires_grad_(True)
        b = torch.randn(shape[1], generator=g).requires_grad_(True) if has_b else None
        y = ref.bias_act.bias_act(x, b, dim=1, act=act, gain=gain, clamp=clamp, impl='ref')
        dy = torch.randn(shape, generator=g)
        grads = torch.autograd.grad(y, [x] + ([b] if has_b else []), dy, create_graph=True)
        # second order: d/d(dy) of <dx, v> = grad-1 pass of v (activation has zero 2nd derivative for linear / lrelu)
        out[f'bias_act/{name}/x'] = npy(x)
        if has_b:
            out[f'bias_act/{name}/b'] = npy(b)
            out[f'bias_act/{name}/db'] = npy(grads[1])
        out[f'bias_act/{name}/y'] = npy(y)
        out[f'bias_act/{name}/dy'] = npy(dy)
        out[f'bias_act/{name}/dx'] = npy(grads[0])
    for act in ALL_ACTS:
        x = torch.randn(3, 5, 4, generator=g, dtype=torch.float64).requires_grad_(True)
        b = torch.randn(5, generator=g, dtype=torch.float64)
        y = ref.bias_act.bias_act(x, b, dim=1, act=act, impl='ref')
        dy = torch.randn(3, 5, 4, generator=g, dtype=torch.float64).requires_grad_(True)
        dx, = torch.autograd.grad(y, x, dy, create_graph=True)
        v = torch.randn(3, 5, 4, generator=g, dtype=torch.float64)
        d_x, d_dy = torch.autograd.grad(dx, [x, dy], v, allow_unused=True)
        out[f'bias_act_all/{act}/x'] = npy(x)
        out[f'bias_act_all/{act}/b'] = npy(b)
        out[f'bias_act_all/{act}/y'] = npy(y)
        out[f'bias_act_all/{act}/dy'] = npy(dy)
        out[f'bias_act_all/{act}/dx'] = npy(dx)
        out[f'bias_act_all/{act}/v'] = npy(v)
        out[f'bias_act_all/{act}/d_dy'] = npy(d_dy)
        out[f'bias_act_all/{act}/d_x'] = npy(d_x) if d_x is not None else np.zeros(x.shape)


def gen_upfirdn2d(ref, out):
    g = torch.Generator().manual_seed(2)
    for name, f, kw, shape in UPFIRDN_CASES:
        ft = ref.upfirdn2d.setup_filter(f) if f is not None else None
        x = torch.randn(shape, generator=g).requires_grad_(True)
        y = ref.upfirdn2d.upfirdn2d(x, ft, impl='ref', **kw)
        dy = torch.randn(y.shape, generator=g)
        dx, = torch.autograd.grad(y, x, dy)
        out[f'upfirdn2d/{name}/x'] = npy(x)
        out[f'upfirdn2d/{name}/f'] = npy(ft) if ft is not None else np.zeros([0], dtype=np.float32)
        out[f'upfirdn2d/{name}/y'] = npy(y)
        out[f'upfirdn2d/{name}/dy'] = npy(dy)
        out[f'upfirdn2d/{name}/dx'] = npy(dx)
    x = torch.randn(2, 2, 6, 7, generator=g)
    f4 = ref.upfirdn2d.setup_filter(F4)
    out['upfirdn2d/wrappers/x'] = npy(x)
    out['upfirdn2d/wrappers/filter2d'] = npy(ref.upfirdn2d.filter2d(x, f4, impl='ref'))
    out['upfirdn2d/wrappers/upsample2d'] = npy(ref.upfirdn2d.upsample2d(x, f4, impl='ref'))
    out['upfirdn2d/wrappers/downsample2d'] = npy(ref.upfirdn2d.downsample2d(x[:, :, :, :6], f4, impl='ref'))
    out['upfirdn2d/setup/f4'] = npy(f4)
    out['upfirdn2d/setup/sym6'] = npy(ref.upfirdn2d.setup_filter(SYM6))
    out['upfirdn2d/setup/f4_flip_gain'] = npy(ref.upfirdn2d.setup_filter([1, 2, 3, 4], flip_filter=True, gain=3))
    out['upfirdn2d/setup/sep'] = npy(ref.upfirdn2d.setup_filter([1, 2, 3], separable=True, gain=2))


def gen_conv(ref, out):
    g = torch.Generator().manual_seed(3)
    f4 = ref.upfirdn2d.setup_filter(F4)
    for name, ci, co, k, kw, h in CONV_RESAMPLE_CASES:
        x = torch.randn(2, ci, h, h, generator=g).requires_grad_(True)
        w = torch.randn(co, ci, k, k, generator=g).requires_grad_(True)
        y = ref.conv2d_resample.conv2d_resample(x, w, f=f4, **kw)
        dy = torch.randn(y.shape, generator=g)
        dx, dw = torch.autograd.grad(y, [x, w], dy)
        for key, val in dict(x=x, w=w, y=y, dy=dy, dx=dx, dw=dw).items():
            out[f'conv2d_resample/{name}/{key}'] = npy(val)
    # modulated conv: both branches, with / without demodulation and noise, up and plain
    for name, kw in [('plain_demod_noise', dict(padding=1)), ('up_demod_noise', dict(up=2, padding=1, flip_weight=False)),
                     ('torgb', dict(demodulate=False, k=1, noise=False))]:
        k = kw.pop('k', 3)
        use_noise = kw.pop('noise', True)
        x = torch.randn(3, 4, 8, 8, generator=g).requires_grad_(True)
        w = torch.randn(6, 4, k, k, generator=g).requires_grad_(True)
        s = (torch.randn(3, 4, generator=g) + 1).requires_grad_(True)
        oh = 16 if kw.get('up', 1) == 2 else 8
        noise = torch.randn(3, 1, oh, oh, generator=g) if use_noise else None
        for fused in (False, True):
            y = ref.networks.modulated_conv2d(x=x, weight=w, styles=s, noise=(noise.clone() if noise is not None else None),
                                              resample_filter=f4, fused_modconv=fused, **kw)
            dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
            dx, dw, ds = torch.autograd.grad(y, [x, w, s], dy)
            tag = f'modconv/{name}/{"fused" if fused else "unfused"}'
            for key, val in dict(y=y, dx=dx, dw=dw, ds=ds).items():
                out[f'{tag}/{key}'] = npy(val)
        for key, val in dict(x=x, w=w, s=s, dy=dy).items():
            out[f'modconv/{name}/{key}'] = npy(val)
        if noise is not None:
            out[f'modconv/{name}/noise'] = npy(noise)
    # fma and grid_sample
    a = torch.randn(2, 3, 4, 4, generator=g)
    b = torch.randn(2, 3, 1, 1, generator=g)
    c = torch.randn(2, 1, 4, 4, generator=g)
    out['fma/a'], out['fma/b'], out['fma/c'] = npy(a), npy(b), npy(c)
    out['fma/y'] = npy(ref.fma.fma(a, b, c))
    img = torch.randn(2, 1, 9, 11, generator=g)
    theta = torch.tensor([[[0.9, 0.2, 0.1], [-0.15, 1.1, -0.05]], [[1.2, 0.0, 0.3], [0.1, 0.8, 0.6]]])
    grid = torch.nn.functional.affine_grid(theta, [2, 1, 7, 8], align_corners=False)
    out['grid_sample/img'], out['grid_sample/grid'] = npy(img), npy(grid)
    out['grid_sample/y'] = npy(ref.grid_sample_gradfix.grid_sample(img, grid))


def gen_augment(ref, out):
    g = torch.Generator().manual_seed(4)
    img = torch.rand(4, 1, 32, 32, generator=g) * 2 - 1
    rgb = torch.rand(2, 3, 16, 16, generator=g) * 2 - 1
    out['augment/img'] = npy(img)
    out['augment/rgb'] = npy(rgb)
    for tag, kw, x in [('claro', AUG_KW, img), ('full', AUG_KW_FULL, img), ('full_rgb', AUG_KW_FULL, rgb)]:
        pipe = ref.augment.AugmentPipe(run_dir=None, batch_size=x.shape[0], **kw)
        pipe.p.copy_(torch.as_tensor(0.7))
        for pct in (0.1, 0.5, 0.9):
            torch.manual_seed(11)     # noise / cutout still draw random numbers in debug mode
            out[f'augment/{tag}/pct{pct}'] = npy(pipe(x, False, debug_percentile=pct))
        torch.manual_seed(123)
        out[f'augment/{tag}/seed123'] = npy(pipe(x, False))
    # gradient through the pipe, first and second order (what R1 needs)
    pipe = ref.augment.AugmentPipe(run_dir=None, batch_size=4, **AUG_KW)
    pipe.p.copy_(torch.as_tensor(1.0))
    x = img.clone().requires_grad_(True)
    torch.manual_seed(5)
    y = pipe(x, False)
    wgt = torch.randn(y.shape, generator=g)
    gx, = torch.autograd.grad((y * wgt).sum(), x, create_graph=True)
    ggx, = torch.autograd.grad(gx.square().sum(), x)
    out['augment/grad/w'] = npy(wgt)
    out['augment/grad/gx'] = npy(gx)
    out['augment/grad/ggx'] = npy(ggx)


def gen_model(ref, out):
    torch.manual_seed(0)
    G = ref.networks.Generator(**G_KW).train().requires_grad_(False)
    D = ref.networks.Discriminator(**D_KW).train().requires_grad_(False)
    # non-trivial values for parameters that initialise to zero, so that their code paths are exercised
    gen = torch.Generator().manual_seed(6)
    for name, p in list(G.named_parameters()) + list(D.named_parameters()):
        if name.endswith('noise_strength'):
            p.copy_(torch.randn([], generator=gen) * 0.1)
        elif name.endswith('bias') and 'affine' not in name:
            p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    for k, v in G.state_dict().items():
        out[f'model/G/{k}'] = npy(v)
    for k, v in D.state_dict().items():
        out[f'model/D/{k}'] = npy(v)
    z = torch.randn(4, G_KW['z_dim'], generator=gen)
    c = torch.nn.functional.one_hot(torch.tensor([0, 1, 1, 0]), 2).float()
    real = torch.rand(4, 1, 32, 32, generator=gen) * 2 - 1
    out['model/z'], out['model/c'], out['model/real'] = npy(z), npy(c), npy(real)
    G.eval()
    out['model/G_eval_const'] = npy(G(z, c, noise_mode='const'))                      # fused (grouped) branch
    G.train()
    out['model/G_train_const'] = npy(G(z, c, noise_mode='const'))                     # non-fused branch
    torch.manual_seed(21)
    out['model/G_train_random'] = npy(G(z, c))
    out['model/D_real'] = npy(D(real, c))

    # The four loss phases, with the ADA pipe Gan-track configures, fixed seeds.
    aug = ref.augment.AugmentPipe(run_dir=None, batch_size=4, **AUG_KW).train().requires_grad_(False)
    aug.p.copy_(torch.as_tensor(0.6))
    loss = ref.loss.StyleGAN2Loss(device=torch.device('cpu'), G=G, D=D, augment_pipe=aug, **LOSS_KW)
    loss.pl_mean.copy_(torch.as_tensor(0.37))
    for phase, module, gain in [('Gmain', G, 1), ('Greg', G, 4), ('Dmain', D, 1), ('Dreg', D, 16)]:
        module.requires_grad_(True)
        for p in module.parameters():
            p.grad = None
        torch.manual_seed(100)
        loss.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, gain=gain, cur_nimg=0)
        module.requires_grad_(False)
        for name, p in module.named_parameters():
            if p.grad is not None:
                out[f'loss/{phase}/{name}'] = npy(p.grad)
        out[f'loss/{phase}/pl_mean'] = npy(loss.pl_mean)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = import_reference()
    torch.set_num_threads(4)
    for fname, fn in [('ops_bias_act.npz', gen_bias_act), ('ops_upfirdn2d.npz', gen_upfirdn2d), ('ops_conv.npz', gen_conv),
                      ('augment.npz', gen_augment), ('model.npz', gen_model)]:
        out = {}
        with torch.no_grad() if False else torch.enable_grad():
            fn(ref, out)
        path = os.path.join(OUT, fname)
        np.savez_compressed(path, **out)
        print(f'{fname}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
