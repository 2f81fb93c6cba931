"""GPU: the fused ADA warp kernel (csrc/augment_warp.cu) against the op-by-op sequence of the reference
(S3/training/augment_mi.py:286-321: reflect pad by host-read margins, upsample2d, affine_grid + grid_sample,
downsample2d) on the same random transforms: values, first-order gradient and the R1-style double backward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[True, False], ids=['two_pass', 'single_gather'], autouse=True)
def warp_mode(request):
    """Every test runs with both forms of the fused warp: two passes through a workspace (default) and the single gather."""
    from gan_track_b200.torch_utils.ops import aug_warp
    old = aug_warp.two_pass
    aug_warp.two_pass = request.param
    yield request.param
    aug_warp.two_pass = old


def _pipe(**kw):
    from gan_track_b200.training import augment
    base = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1)
    base.update(kw)
    pipe = augment.AugmentPipe(**base).cuda()
    pipe.p.fill_(1.0)
    return pipe


def _run(pipe, x, fused, seed):
    pipe.fused_warp = fused
    torch.manual_seed(seed)
    return pipe(x)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize('shape', [(8, 1, 64, 64), (4, 3, 48, 80), (6, 1, 256, 256)])
@pytest.mark.parametrize('cfg', ['claro', 'full'])
def test_fused_warp_matches_reference_sequence(shape, cfg):
    kw = dict(xint_max=0.05, rotate_max=3 / 360, scale_std=0.05, aniso_std=0.05, xfrac_std=0.05, rotate90=0) if cfg == 'claro' else {}
    pipe = _pipe(**kw)
    x = torch.randn(shape, device='cuda')
    for seed in range(3):
        y_ref = _run(pipe, x, False, seed)
        y = _run(pipe, x, True, seed)
        assert y.shape == y_ref.shape
        # white-noise input is the worst case: the two formulations round the fp32 sample coordinates differently (~3e-5 px at
        # 512-px extents) and noise has unit slope per pixel; smooth images are checked tighter below
        assert _rel(y, y_ref) <= 5e-5, f'seed {seed}'


def test_fused_warp_smooth_image_tight():
    pipe = _pipe(xint_max=0.05, rotate_max=3 / 360, scale_std=0.05, aniso_std=0.05, xfrac_std=0.05, rotate90=0)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 256, device='cuda'), torch.linspace(-1, 1, 256, device='cuda'), indexing='ij')
    x = (torch.sin(6 * xx) * torch.cos(4 * yy) + 0.3 * xx * yy)[None, None].repeat(4, 1, 1, 1).contiguous()
    for seed in range(3):
        assert _rel(_run(pipe, x, True, seed), _run(pipe, x, False, seed)) <= 1e-5


def test_fused_warp_gradients_and_double_backward():
    pipe = _pipe()
    x0 = torch.randn([4, 1, 64, 64], device='cuda')
    w = torch.randn([4, 1, 64, 64], device='cuda')
    outs = {}
    for fused in (False, True):
        x = x0.clone().requires_grad_(True)
        y = _run(pipe, x, fused, 7)
        g, = torch.autograd.grad((y * w).sum() + (y.square()).sum(), x, create_graph=True)      # nonlinear head -> g depends on x
        pen = g.square().sum()
        gg, = torch.autograd.grad(pen, x)
        outs[fused] = (y.detach(), g.detach(), gg.detach())
    for a, b, name in zip(outs[True], outs[False], ['y', 'dx', 'd(|dx|^2)/dx']):
        assert _rel(a, b) <= 5e-5, name


def test_fused_warp_has_no_host_sync():
    """The fused path must be capturable: run it under a CUDA-graph capture."""
    pipe = _pipe()
    x = torch.randn([4, 1, 64, 64], device='cuda')
    pipe.fused_warp = True
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            pipe(x)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = pipe(x)
    g.replay()
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()


def test_identity_transform_returns_the_image():
    pipe = _pipe()
    pipe.p.fill_(0.0)        # every gate closed -> G_inv = I; the 2x up / down filters are a half-band pair
    x = torch.randn([2, 1, 64, 64], device='cuda')
    y = _run(pipe, x, True, 0)
    y_ref = _run(pipe, x, False, 0)
    assert _rel(y, y_ref) <= 2e-5


@pytest.mark.parametrize('kw_name', ['claro', 'all_geometric'])
def test_fused_parameter_kernel_equals_op_chain(kw_name):
    """csrc/augment_params.cu (gates, 3x3 matrix chain, margins, sampling matrix in one launch) against the pipe's own op-by-op
    parameter algebra on the SAME random draws (same seed, same draw order): identical margins, images equal to fp32 rounding --
    for Gan-track's transform set and for every geometric transform at the upstream default ranges, at several strengths."""
    from oracle import gen_golden as gg
    from gan_track_b200.training import augment
    kw = gg.AUG_KW if kw_name == 'claro' else dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1)
    pipe = augment.AugmentPipe(**kw).cuda()
    g = torch.Generator().manual_seed(3)
    x = (torch.rand([16, 1, 64, 64], generator=g) * 2 - 1).cuda()
    for p in (0.0, 0.3, 1.0):
        pipe.p.copy_(torch.as_tensor(p))
        outs = []
        for fused in (True, False):
            pipe.fused_params = fused
            torch.manual_seed(1234)
            outs.append(pipe(x, False))
        torch.manual_seed(99)
        after_fused = torch.rand([4], device='cuda')            # both routes must leave the generator in the same state
        err = float((outs[0] - outs[1]).abs().max() / outs[1].abs().max())
        assert err <= 2e-5, (kw_name, p, err)
    pipe.fused_params = True
    # gradient flows through the warp exactly as before (parameters carry no gradient)
    xr = x.clone().requires_grad_(True)
    torch.manual_seed(5)
    y = pipe(xr, False)
    gx, = torch.autograd.grad(y.square().sum(), xr)
    assert torch.isfinite(gx).all() and float(gx.abs().max()) > 0
    del after_fused
