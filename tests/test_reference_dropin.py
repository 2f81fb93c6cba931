"""INTEGRATION.md option A, executed: the reference's UNMODIFIED model / loss files (S3/training/networks_stylegan2.py,
S3/training/loss.py) run over this repo's `torch_utils.ops` package in place of their own -- `from torch_utils.ops import
conv2d_resample, upfirdn2d, bias_act, fma` resolves to gan_track_b200's modules -- and reproduce the committed golden outputs
(which are the reference's outputs over ITS ops, oracle/gen_golden.py).  That pins the op surface the reference's callers see:
module names, function names, keyword arguments, defaults, `conv2d_gradfix.no_weight_gradients`, `activation_funcs`.

The reference tree only exists in the build container (never on the GPU box), so this is a CPU test that is skipped without it;
the primitive ops are the oracle's (the product has no CPU path), exactly as in tests/test_host_golden.py.  It runs in a
subprocess because the reference's top-level package names (`torch_utils`, `training`, `dnnlib`) must not leak into this
interpreter.
"""
import os
import subprocess
import sys

import pytest

S3 = '/root/reference/src/models/stylegan3'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import importlib, sys, types
import numpy as np, torch
ROOT, S3 = sys.argv[1], sys.argv[2]
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(1, S3)

import gan_track_b200.torch_utils.ops as our_ops
from gan_track_b200.torch_utils.ops import bias_act, conv2d_gradfix, conv2d_resample, fma, grid_sample_gradfix, upfirdn2d
from oracle.backend import oracle_ops
from oracle import gen_golden as gg

import torch_utils                                   # the REFERENCE's package (misc, persistence stay the reference's own)
assert torch_utils.__file__.startswith(S3)
ops_pkg = types.ModuleType('torch_utils.ops')        # ... with its op sub-package replaced by this repo's modules
ops_pkg.__path__ = []
for name, mod in dict(bias_act=bias_act, conv2d_gradfix=conv2d_gradfix, conv2d_resample=conv2d_resample, fma=fma,
                      grid_sample_gradfix=grid_sample_gradfix, upfirdn2d=upfirdn2d).items():
    setattr(ops_pkg, name, mod)
    sys.modules['torch_utils.ops.' + name] = mod
sys.modules['torch_utils.ops'] = ops_pkg
torch_utils.ops = ops_pkg

from training import networks_stylegan2 as ref_nets  # unmodified reference files
from training import loss as ref_loss
for m in ['matplotlib', 'matplotlib.pyplot', 'openpyxl']:      # plotting / spreadsheet imports of augment_mi, unused here
    sys.modules.setdefault(m, types.ModuleType(m))
from training import augment_mi as ref_aug
assert ref_nets.__file__.startswith(S3) and ref_loss.__file__.startswith(S3)
assert ref_nets.bias_act is bias_act and ref_nets.conv2d_resample is conv2d_resample and ref_nets.upfirdn2d is upfirdn2d and ref_nets.fma is fma
assert ref_loss.conv2d_gradfix is conv2d_gradfix and ref_loss.upfirdn2d is upfirdn2d
assert ref_aug.__file__.startswith(S3) and ref_aug.grid_sample_gradfix is grid_sample_gradfix and ref_aug.upfirdn2d is upfirdn2d

Z = np.load(ROOT + '/tests/golden/model.npz')
t = lambda k: torch.from_numpy(Z[k])


def load(module, prefix):
    module.load_state_dict({k[len(prefix):]: t(k) for k in Z.files if k.startswith(prefix)}, strict=True)
    return module


def close(a, key, rtol, atol):
    b = t(key)
    err = float((a.detach() - b).abs().max())
    assert err <= atol + rtol * float(b.abs().max()), (key, err)
    return err


G = load(ref_nets.Generator(**gg.G_KW), 'model/G/').train().requires_grad_(False)
D = load(ref_nets.Discriminator(**gg.D_KW), 'model/D/').train().requires_grad_(False)
z, c, real = t('model/z'), t('model/c'), t('model/real')
with oracle_ops():
    G.eval()
    e0 = close(G(z, c, noise_mode='const'), 'model/G_eval_const', 1e-4, 2e-5)
    G.train()
    e1 = close(G(z, c, noise_mode='const'), 'model/G_train_const', 1e-4, 2e-5)
    e2 = close(D(real, c), 'model/D_real', 1e-4, 2e-5)
    # all four loss phases (both regularisers' double backwards) through the reference's own StyleGAN2Loss and AugmentPipe
    aug = ref_aug.AugmentPipe(run_dir=None, batch_size=real.shape[0], **gg.AUG_KW).train().requires_grad_(False)
    aug.p.copy_(torch.as_tensor(0.6))
    L = ref_loss.StyleGAN2Loss(device=torch.device('cpu'), G=G, D=D, augment_pipe=aug, **gg.LOSS_KW)
    for phase, module, gain in [('Gmain', G, 1), ('Greg', G, 4), ('Dmain', D, 1), ('Dreg', D, 16)]:
        L.pl_mean.copy_(torch.as_tensor(0.37))
        module.requires_grad_(True)
        for p in module.parameters():
            p.grad = None
        torch.manual_seed(100)
        L.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, gain=gain, cur_nimg=0)
        module.requires_grad_(False)
        checked, worst = 0, 0.0
        for name, p in module.named_parameters():
            key = 'loss/%s/%s' % (phase, name)
            if key not in Z.files:
                assert p.grad is None, (phase, name)
                continue
            ref = Z[key]
            if p.grad is None:
                assert not np.any(ref), (phase, name)      # see tests/test_host_golden.py: exact zeros of the CPU ref path
                continue
            scale = max(float(np.abs(ref).max()), 1e-6)
            np.testing.assert_allclose(p.grad.numpy(), ref, rtol=2e-3, atol=2e-4 * scale, err_msg=phase + ' ' + name)
            worst = max(worst, float(np.abs(p.grad.numpy() - ref).max()) / scale)
            checked += 1
        assert checked > 10, (phase, checked)
        print('%s: %d gradient tensors, worst error %.2g of the tensor maximum' % (phase, checked, worst))
print('dropin ok', e0, e1, e2)
'''


@pytest.mark.skipif(not os.path.isdir(S3), reason='the reference tree exists only in the build container')
def test_reference_model_files_run_unmodified_over_this_op_package():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, '-c', CHILD, ROOT, S3], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert 'dropin ok' in r.stdout and 'Dreg:' in r.stdout, r.stdout[-2000:]
