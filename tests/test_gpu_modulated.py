"""GPU: fused modulation / demodulation+noise+bias+activation kernels (csrc/modulated.cu) against the reference's op
sequence (S3/training/networks_stylegan2.py:69-72, 325-327: multiply, fma, bias_act) built from this package's
individually verified ops: values, first-order gradients of every input, and second-order gradients (the path-length
regulariser differentiates the backward pass)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _ref_mod_scale(x, s):
    return x * s.to(x.dtype).reshape(x.shape[0], -1, 1, 1)


def _ref_demod_act(x, d, noise, b, act, gain, clamp):
    from gan_track_b200.torch_utils.ops import bias_act, fma
    if d is not None and noise is not None:
        x = fma.fma(x, d.to(x.dtype).reshape(x.shape[0], -1, 1, 1), noise.to(x.dtype))
    elif d is not None:
        x = x * d.to(x.dtype).reshape(x.shape[0], -1, 1, 1)
    elif noise is not None:
        x = x + noise.to(x.dtype)
    return bias_act.bias_act(x, b.to(x.dtype) if b is not None else None, act=act, gain=gain, clamp=clamp)


SHAPES = [(4, 64, 16, 16), (2, 512, 8, 8), (3, 128, 9, 7), (2, 256, 5, 5),
          # >= 1 MB per sample: the bulk-copy staged kernels (csrc/stream_bulk.cuh), incl. a short last chunk and C/VEC > 32
          (2, 64, 128, 128), (2, 128, 64, 72), (1, 256, 64, 64), (1, 512, 40, 32)]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('shape', SHAPES)
def test_mod_scale(shape, dtype):
    from gan_track_b200.torch_utils.ops import modulated
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    N, C, H, W = shape
    x0 = _cl(torch.randn(shape, device='cuda').to(dtype))
    s0 = torch.randn([N, C], device='cuda')
    gy = _cl(torch.randn(shape, device='cuda').to(dtype))
    assert modulated.applicable(x0)
    outs = []
    for fn in (modulated.mod_scale, _ref_mod_scale):
        x, s = x0.clone().requires_grad_(True), s0.clone().requires_grad_(True)
        y = fn(x, s)
        gx, gs = torch.autograd.grad(y, [x, s], gy, create_graph=True)
        # second order: a scalar of the first-order gradients, differentiated w.r.t. x, s (what Greg does through ws)
        q = (gs.float().square().sum() + (gx.float() * x.float()).sum())
        ggx, ggs = torch.autograd.grad(q, [x, s])
        outs.append((y, gx, gs, ggx, ggs))
    for a, b, name in zip(outs[0], outs[1], ['y', 'gx', 'gs', 'd2/dx', 'd2/ds']):
        assert _rel(a, b) <= (tol if 'd2' not in name else 5 * tol), name


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('shape', SHAPES)
@pytest.mark.parametrize('cfg', [('lrelu', np.sqrt(2), 256.0, True, True), ('lrelu', 1.0, 1.5, True, True), ('linear', 1.0, None, False, True),
                                 ('lrelu', np.sqrt(2), None, True, False)], ids=['lrelu_c256', 'lrelu_tightclamp', 'linear_nodemod', 'lrelu_nonoise'])
def test_demod_act(shape, dtype, cfg):
    from gan_track_b200.torch_utils.ops import modulated
    act, gain, clamp, demod, with_noise = cfg
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    N, C, H, W = shape
    x0 = _cl(torch.randn(shape, device='cuda').to(dtype))
    d0 = torch.rand([N, C], device='cuda') + 0.5 if demod else None
    nz0 = (torch.randn([N, 1, H, W], device='cuda') * 0.3) if with_noise else None
    b0 = torch.randn([C], device='cuda').to(dtype)
    gy = _cl(torch.randn(shape, device='cuda').to(dtype))
    outs = []
    for fused in (True, False):
        x = x0.clone().requires_grad_(True)
        d = d0.clone().requires_grad_(True) if demod else None
        nz = nz0.clone().requires_grad_(True) if with_noise else None
        b = b0.clone().requires_grad_(True)
        if fused:
            y = modulated.demod_act(x, d, nz, b, act=act, alpha=0.2, gain=gain, clamp=clamp)
        else:
            y = _ref_demod_act(x, d, nz, b, act, gain, clamp)
        ins = [t for t in (x, d, nz, b) if t is not None]
        grads = torch.autograd.grad(y, ins, gy, create_graph=True)
        q = sum((g.float() * torch.arange(1, g.numel() + 1, device='cuda').reshape(g.shape).float().remainder(7).add(1)).sum() * 1e-2 for g in grads)
        q = q + grads[0].float().square().sum() * 1e-2
        second = torch.autograd.grad(q, [t for t in (x, d) if t is not None], allow_unused=True) if (demod and q.requires_grad) else ()
        outs.append((y, *grads, *[g for g in second if g is not None]))
    assert len(outs[0]) == len(outs[1])
    for i, (a, b_) in enumerate(zip(outs[0], outs[1])):
        assert a.shape == b_.shape
        assert _rel(a, b_) <= 5 * tol, f'output {i}'


def test_synthesis_layer_fused_vs_unfused_vs_fp32():
    """A whole SynthesisLayer (up=1 and up=2) in fp16: the fused element-wise route and the reference op sequence are both
    compared with the same layer evaluated in fp32; the fused route must be within the fp16 tolerance of the north star
    (1e-2) or at least as close to fp32 as the unfused sequence is."""
    from gan_track_b200.torch_utils.ops import modulated
    from gan_track_b200.training import networks_stylegan2 as nets
    torch.manual_seed(0)
    for up in (1, 2):
        layer = nets.SynthesisLayer(64, 128, w_dim=32, resolution=32, up=up, conv_clamp=256, channels_last=True).cuda()
        layer.noise_strength.data.fill_(0.3)
        x0 = torch.randn([3, 64, 32 // up, 32 // up], device='cuda')
        w0 = torch.randn([3, 32], device='cuda')
        res = {}
        for mode in ('fp32', 'fused', 'unfused'):
            saved = modulated.applicable
            if mode != 'fused':
                modulated.applicable = lambda t: False
            try:
                x = (x0.clone() if mode == 'fp32' else _cl(x0.half())).requires_grad_(True)
                w = w0.clone().requires_grad_(True)
                torch.manual_seed(1)
                y = layer(x, w, noise_mode='random', fused_modconv=False)
                g = torch.autograd.grad(y.float().square().sum(), [x, w, layer.weight, layer.bias, layer.noise_strength, layer.affine.weight])
            finally:
                modulated.applicable = saved
            res[mode] = (y, *g)
        for i, (t, a, b_) in enumerate(zip(res['fp32'], res['fused'], res['unfused'])):
            ef, eu = _rel(a, t), _rel(b_, t)
            assert ef <= max(1e-2, 1.25 * eu), f'up={up} output {i}: fused {ef:.3e} unfused {eu:.3e}'
