"""CPU: the oracle restatements (oracle/ops_ref.py) against the golden vectors produced by the real reference
(oracle/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import gen_golden as gg
from oracle import ops_ref as R


def t(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize('case', gg.BIAS_ACT_PATH_CASES, ids=lambda c: c[0])
def test_bias_act_path_cases(golden, case):
    name, act, gain, clamp, has_b, shape = case
    G = golden('ops_bias_act.npz')
    x = t(G[f'bias_act/{name}/x']).requires_grad_(True)
    b = t(G[f'bias_act/{name}/b']).requires_grad_(True) if has_b else None
    y = R.bias_act(x, b, dim=1, act=act, gain=gain, clamp=clamp)
    close(y, G[f'bias_act/{name}/y'])
    dy = t(G[f'bias_act/{name}/dy'])
    grads = torch.autograd.grad(y, [x] + ([b] if has_b else []), dy)
    close(grads[0], G[f'bias_act/{name}/dx'])
    if has_b:
        close(grads[1], G[f'bias_act/{name}/db'], rtol=1e-4, atol=1e-5)
    # the kernel-semantics restatement must agree with autograd of the ref path for the saved-output activations
    _, _, _, refmode, _ = R.ACTIVATIONS[act]
    alpha, g_, c_ = R._resolve(act, None, gain, clamp)
    yk = R.bias_act_kernel(x.detach(), b.detach() if has_b else None, None, None, None, 0, 1, act, alpha, g_, c_)
    close(yk, G[f'bias_act/{name}/y'])
    if 'y' in refmode:
        dxk = R.bias_act_kernel(dy, None, None, y.detach(), None, 1, 1, act, alpha, g_, c_)
        close(dxk, G[f'bias_act/{name}/dx'])


@pytest.mark.parametrize('act', gg.ALL_ACTS)
def test_bias_act_all_activations_fp64(golden, act):
    G = golden('ops_bias_act.npz')
    x = t(G[f'bias_act_all/{act}/x']).requires_grad_(True)
    b = t(G[f'bias_act_all/{act}/b'])
    dy = t(G[f'bias_act_all/{act}/dy']).requires_grad_(True)
    v = t(G[f'bias_act_all/{act}/v'])
    y = R.bias_act(x, b, dim=1, act=act)
    close(y, G[f'bias_act_all/{act}/y'], 1e-12, 1e-12)
    dx, = torch.autograd.grad(y, x, dy, create_graph=True)
    close(dx, G[f'bias_act_all/{act}/dx'], 1e-10, 1e-12)
    # kernel semantics, all three grad levels, against the reference's autograd
    _, alpha, gain, refmode, has2 = R.ACTIVATIONS[act]
    xr = x.detach() if ('x' in refmode or has2) else None
    br = b if ('x' in refmode or has2) else None
    yr = y.detach() if 'y' in refmode else None
    close(R.bias_act_kernel(x.detach(), b, None, None, None, 0, 1, act, alpha, gain, -1.0), G[f'bias_act_all/{act}/y'], 1e-10, 1e-12)
    close(R.bias_act_kernel(dy.detach(), br, xr, yr, None, 1, 1, act, alpha, gain, -1.0), G[f'bias_act_all/{act}/dx'], 1e-8, 1e-10)
    close(R.bias_act_kernel(v, br, xr, yr, None, 1, 1, act, alpha, gain, -1.0), G[f'bias_act_all/{act}/d_dy'], 1e-8, 1e-10)
    if has2:
        close(R.bias_act_kernel(v, br, xr, yr, dy.detach(), 2, 1, act, alpha, gain, -1.0), G[f'bias_act_all/{act}/d_x'], 1e-7, 1e-9)


@pytest.mark.parametrize('case', gg.UPFIRDN_CASES, ids=lambda c: c[0])
def test_upfirdn2d(golden, case):
    name, f, kw, shape = case
    G = golden('ops_upfirdn2d.npz')
    x = t(G[f'upfirdn2d/{name}/x']).requires_grad_(True)
    ft = R.setup_filter(f) if f is not None else None
    if ft is not None:
        close(ft, G[f'upfirdn2d/{name}/f'], 1e-7, 0)
    y = R.upfirdn2d(x, ft, **kw)
    close(y, G[f'upfirdn2d/{name}/y'])
    dy = t(G[f'upfirdn2d/{name}/dy'])
    dx, = torch.autograd.grad(y, x, dy)
    close(dx, G[f'upfirdn2d/{name}/dx'])
    # independent direct-form statement, and the backward-is-the-same-op rule
    close(R.upfirdn2d_direct(x.detach(), ft, **kw), G[f'upfirdn2d/{name}/y'], 1e-4, 1e-5)
    bk = R.upfirdn2d_backward_args(x.shape, y.shape, ft, kw.get('up', 1), kw.get('down', 1), kw.get('padding', 0),
                                   kw.get('flip_filter', False), kw.get('gain', 1))
    close(R.upfirdn2d(dy, ft, **bk), G[f'upfirdn2d/{name}/dx'], 1e-4, 1e-5)


def test_upfirdn2d_wrappers_and_filters(golden):
    G = golden('ops_upfirdn2d.npz')
    x = t(G['upfirdn2d/wrappers/x'])
    f4 = R.setup_filter(gg.F4)
    close(R.filter2d(x, f4), G['upfirdn2d/wrappers/filter2d'])
    close(R.upsample2d(x, f4), G['upfirdn2d/wrappers/upsample2d'])
    close(R.downsample2d(x[:, :, :, :6], f4), G['upfirdn2d/wrappers/downsample2d'])
    close(f4, G['upfirdn2d/setup/f4'], 1e-7, 0)
    close(R.setup_filter(gg.SYM6), G['upfirdn2d/setup/sym6'], 1e-7, 0)
    close(R.setup_filter([1, 2, 3, 4], flip_filter=True, gain=3), G['upfirdn2d/setup/f4_flip_gain'], 1e-7, 0)
    close(R.setup_filter([1, 2, 3], separable=True, gain=2), G['upfirdn2d/setup/sep'], 1e-7, 0)
    # DC gain of upsample2d is 1
    ones = torch.ones(1, 1, 8, 8)
    assert torch.allclose(R.upsample2d(ones, f4)[:, :, 3:-3, 3:-3], torch.ones(1, 1, 10, 10), atol=1e-6)


@pytest.mark.parametrize('case', gg.CONV_RESAMPLE_CASES, ids=lambda c: c[0])
def test_conv2d_resample(golden, case):
    name, ci, co, k, kw, h = case
    G = golden('ops_conv.npz')
    x = t(G[f'conv2d_resample/{name}/x']).requires_grad_(True)
    w = t(G[f'conv2d_resample/{name}/w']).requires_grad_(True)
    y = R.conv2d_resample(x, w, f=R.setup_filter(gg.F4), **kw)
    close(y, G[f'conv2d_resample/{name}/y'], 1e-5, 1e-5)
    dx, dw = torch.autograd.grad(y, [x, w], t(G[f'conv2d_resample/{name}/dy']))
    close(dx, G[f'conv2d_resample/{name}/dx'], 1e-5, 1e-5)
    close(dw, G[f'conv2d_resample/{name}/dw'], 1e-4, 1e-4)


@pytest.mark.parametrize('name,kw', [('plain_demod_noise', dict(padding=1)), ('up_demod_noise', dict(up=2, padding=1, flip_weight=False)),
                                     ('torgb', dict(demodulate=False))])
def test_modulated_conv2d(golden, name, kw):
    G = golden('ops_conv.npz')
    x = t(G[f'modconv/{name}/x']).requires_grad_(True)
    w = t(G[f'modconv/{name}/w']).requires_grad_(True)
    s = t(G[f'modconv/{name}/s']).requires_grad_(True)
    noise = t(G[f'modconv/{name}/noise']) if f'modconv/{name}/noise' in G else None
    dy = t(G[f'modconv/{name}/dy'])
    for fused in (False, True):
        tag = f'modconv/{name}/{"fused" if fused else "unfused"}'
        y = R.modulated_conv2d(x, w, s, noise=noise, resample_filter=R.setup_filter(gg.F4), fused_modconv=fused, **kw)
        close(y, G[f'{tag}/y'], 1e-4, 1e-5)
        dx, dw, ds = torch.autograd.grad(y, [x, w, s], dy)
        close(dx, G[f'{tag}/dx'], 1e-4, 1e-5)
        close(dw, G[f'{tag}/dw'], 1e-4, 1e-4)
        close(ds, G[f'{tag}/ds'], 1e-4, 1e-4)
    if kw.get('demodulate', True):
        d = R.dcoefs_closed_form(w.detach(), s.detach())
        wfull = (w.detach()[None] * s.detach()[:, None, :, None, None]).double()
        close(d, (wfull.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().numpy(), 1e-6, 0)


def test_fma_and_grid_sample(golden):
    G = golden('ops_conv.npz')
    close(R.fma(t(G['fma/a']), t(G['fma/b']), t(G['fma/c'])), G['fma/y'])
    img, grid = t(G['grid_sample/img']), t(G['grid_sample/grid'])
    close(R.grid_sample(img, grid), G['grid_sample/y'])
    close(R.grid_sample_direct(img, grid), G['grid_sample/y'], 1e-4, 1e-5)
