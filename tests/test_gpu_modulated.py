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


def _ref_prep(weight, styles, prenorm):
    """The reference's tensor-op chain (S3/training/networks_stylegan2.py:52-63)."""
    O, I, kh, kw = weight.shape
    w16 = sn = None
    if prenorm:
        weight = weight * (1 / np.sqrt(I * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
        w16, sn = weight.to(torch.float16), styles
    wsq = weight.square().sum(dim=[2, 3])
    return w16, sn, (styles.square() @ wsq.t() + 1e-8).rsqrt()


@pytest.mark.parametrize('N,O,I,k,prenorm', [(32, 512, 512, 3, True), (32, 64, 128, 3, True), (16, 256, 512, 3, False), (8, 512, 512, 3, False),
                                             (5, 72, 36, 1, True), (64, 128, 64, 3, True)])
def test_modprep_matches_tensor_ops(N, O, I, k, prenorm):
    """csrc/modprep.cu (pre-normalisation + demodulation coefficients in four launches) against the reference's op chain: the
    scaled fp16 weight and normalised styles bit for bit, dcoefs to fp32 rounding, and the gradients w.r.t. weight and styles
    for random upstream gradients of all three outputs."""
    from gan_track_b200.torch_utils.ops import modulated
    torch.manual_seed(N + O)
    weight = torch.randn(O, I, k, k, device='cuda').requires_grad_(True)
    styles = (torch.randn(N, I, device='cuda') * 1.5 + 1).requires_grad_(True)
    assert modulated.prep_applicable(weight, styles)
    w16, sn, d = modulated.prep(weight, styles, prenorm)
    rw, rs, rd = _ref_prep(weight, styles, prenorm)
    if prenorm:
        assert torch.equal(w16, rw) and torch.equal(sn, rs)
    else:
        assert w16 is None and sn is None
    assert _rel(d, rd) <= 2e-6
    g_d = torch.randn_like(d)
    outs, routs, gouts = [d], [rd], [g_d]
    if prenorm:
        g_w = (torch.randn_like(w16.float()) * 0.1).half()
        g_s = torch.randn_like(sn)
        outs, routs, gouts = [w16, sn, d], [rw, rs, rd], [g_w, g_s, g_d]
    gW, gs = torch.autograd.grad(outs, [weight, styles], gouts)
    rW, rgs = torch.autograd.grad(routs, [weight, styles], gouts)
    assert _rel(gW, rW) <= 1e-5, _rel(gW, rW)
    assert _rel(gs, rgs) <= 1e-5, _rel(gs, rgs)
    # only dcoefs used downstream (no gradient reaches the other outputs)
    w16, sn, d = modulated.prep(weight, styles, prenorm)
    gW2, gs2 = torch.autograd.grad([d], [weight, styles], [g_d])
    rW2, rgs2 = torch.autograd.grad([_ref_prep(weight, styles, prenorm)[2]], [weight, styles], [g_d])
    assert _rel(gW2, rW2) <= 1e-5 and _rel(gs2, rgs2) <= 1e-5
    # closed under differentiation to second order (the path-length pass differentiates G's backward, S3/training/loss.py:85-100):
    # the style side's backward has a closed-form backward of its own (csrc/modprep.cu style_bwd2_*), checked against autograd's double
    # backward of the tensor-op chain -- with every output used non-linearly, as mod_scale / demod_act use them
    def second_order(prep_fn):
        w16, sn, d = prep_fn(weight, styles, prenorm)
        probe_d = torch.randn(d.shape, device='cuda', generator=torch.Generator('cuda').manual_seed(3))
        y = (d * probe_d).sum() + (d.square() * probe_d.flip(0)).sum()
        if prenorm:
            probe_s = torch.randn(sn.shape, device='cuda', generator=torch.Generator('cuda').manual_seed(4))
            y = y + (sn * probe_s).sum() + (sn.square() * probe_s.flip(1)).sum() + (sn[:, :8].sum(1, keepdim=True) * d).sum()
        g, = torch.autograd.grad([y], [styles], create_graph=True)
        gg_w, gg_s = torch.autograd.grad(g.square().sum(), [weight, styles])
        return g.detach(), gg_w, gg_s
    before, stats0 = _lib_launches(), dict(modulated.prep_stats)
    g, gg_w, gg_s = second_order(modulated.prep)
    used = _lib_launches() - before
    assert modulated.prep_stats['style_bwd2_closed_form'] == stats0['style_bwd2_closed_form'] + 1
    assert modulated.prep_stats['style_bwd2_autograd'] == stats0['style_bwd2_autograd']
    rg, rgg_w, rgg_s = second_order(_ref_prep)
    assert _rel(g, rg) <= 1e-5 and _rel(gg_w, rgg_w) <= 1e-4 and _rel(gg_s, rgg_s) <= 1e-4, (_rel(g, rg), _rel(gg_w, rgg_w), _rel(gg_s, rgg_s))
    assert used <= 24, f'{used} launches: the second-order pass must stay on the fused kernels'
    # a cotangent on the weight-side gradient (double backward w.r.t. the weight) takes the autograd fallback and still matches
    def second_order_w(prep_fn):
        w16, sn, d = prep_fn(weight, styles, prenorm)
        y = (d * torch.linspace(-1, 1, d.numel(), device='cuda').reshape(d.shape)).sum()
        gw, gs_ = torch.autograd.grad([y], [weight, styles], create_graph=True)
        return torch.autograd.grad(gw.square().sum() + gs_.square().sum(), [weight, styles])
    stats0 = dict(modulated.prep_stats)
    a_w, a_s = second_order_w(modulated.prep)
    assert modulated.prep_stats['style_bwd2_autograd'] == stats0['style_bwd2_autograd'] + 1
    b_w, b_s = second_order_w(_ref_prep)
    assert _rel(a_w, b_w) <= 1e-4 and _rel(a_s, b_s) <= 1e-4, (_rel(a_w, b_w), _rel(a_s, b_s))
    # the path-length switch no longer changes the route (it only concerns the fused ToRGB)
    from gan_track_b200.torch_utils.ops import rgb
    with rgb.op_by_op_torgb():
        assert modulated.prep_applicable(weight, styles)


def _lib_launches():
    from gan_track_b200 import _lib
    return _lib.launches


@pytest.mark.parametrize('dtype,C', [(torch.float16, 64), (torch.float32, 32)])
def test_modulated_conv2d_prep_route_equals_op_chain(dtype, C):
    """modulated_conv2d through the fused preparation vs the tensor-op chain (`modulated.fused_prep = False`): output and
    the gradients of x, weight, styles."""
    from gan_track_b200.training import networks_stylegan2 as nets
    torch.manual_seed(7)
    x = _cl(torch.randn(4, C, 32, 32, device='cuda').to(dtype)).requires_grad_(True)
    weight = torch.randn(2 * C, C, 3, 3, device='cuda').requires_grad_(True)
    styles = (torch.randn(4, C, device='cuda') + 1).requires_grad_(True)
    noise = torch.randn(4, 1, 32, 32, device='cuda')
    from gan_track_b200.torch_utils.ops import modulated
    res = []
    for fused in (True, False):
        modulated.fused_prep = fused
        try:
            y = nets.modulated_conv2d(x, weight, styles, noise=noise, padding=1, fused_modconv=False)
        finally:
            modulated.fused_prep = True
        dy = torch.randn(y.shape, device='cuda', generator=torch.Generator('cuda').manual_seed(1)).to(dtype)
        res.append((y.detach(),) + torch.autograd.grad(y, [x, weight, styles], dy))
    tol = 1e-2 if dtype == torch.float16 else 1e-5
    for a, b, name in zip(res[0], res[1], ['y', 'gx', 'gw', 'gs']):
        assert _rel(a, b) <= tol, (name, _rel(a, b))
