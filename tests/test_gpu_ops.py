"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same inputs, against the committed
golden vectors from the real reference, and -- at BASELINE.json's full sizes -- through size-independent properties.
Tolerances are the north star's: fp32 <= 1e-5 relative, fp16 <= 1e-2."""
import numpy as np
import pytest
import torch

from oracle import gen_golden as gg
from oracle import ops_ref as R

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gan_track_b200 import _lib
    from gan_track_b200.torch_utils import custom_ops
    from gan_track_b200.torch_utils.ops import bias_act, conv2d_gradfix, conv2d_resample, fma, grid_sample_gradfix, upfirdn2d

DEV = 'cuda'


def t(a, dtype=None):
    x = torch.from_numpy(np.asarray(a)).to(DEV)
    return x.to(dtype) if dtype is not None else x


def rel_err(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def assert_close(a, b, tol, what=''):
    e = rel_err(a, b)
    assert e <= tol, f'{what}: relative error {e:.3e} > {tol:.1e}'


TOL = {torch.float32: 1e-5, torch.float16: 1e-2, torch.float64: 1e-10}


def test_library_loaded_and_symbols():
    lib = _lib.load()
    assert lib.gt_abi_version() >= 1
    assert lib.gt_sm_count() > 0


# ---------------------------------------------------------------------------------------------------- bias_act

@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('case', gg.BIAS_ACT_PATH_CASES, ids=lambda c: c[0])
def test_bias_act_path_cases(golden, case, cl, dtype):
    name, act, gain, clamp, has_b, shape = case
    if cl and len(shape) != 4:
        pytest.skip('channels-last needs 4-D')
    G = golden('ops_bias_act.npz')
    x = t(G[f'bias_act/{name}/x'], dtype)
    b = t(G[f'bias_act/{name}/b'], dtype) if has_b else None
    dy = t(G[f'bias_act/{name}/dy'], dtype)
    if cl:
        x, dy = x.contiguous(memory_format=torch.channels_last), dy.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    if has_b:
        b.requires_grad_(True)
    y = bias_act.bias_act(x, b, dim=1, act=act, gain=gain, clamp=clamp)
    # oracle on the same (rounded) inputs, in fp32
    xo = x.detach().float().cpu().requires_grad_(True)
    bo = b.detach().float().cpu().requires_grad_(True) if has_b else None
    yo = R.bias_act(xo, bo, dim=1, act=act, gain=gain, clamp=clamp)
    assert_close(y, yo, TOL[dtype], 'y')
    if dtype == torch.float32:
        assert_close(y, G[f'bias_act/{name}/y'], 1e-5, 'y vs golden')
    # first order, differentiable path (create_graph) and fused path (no graph)
    go = torch.autograd.grad(yo, [xo] + ([bo] if has_b else []), dy.float().cpu())
    for create_graph in (True, False):
        g = torch.autograd.grad(y, [x] + ([b] if has_b else []), dy, create_graph=create_graph, retain_graph=True)
        if act == 'linear' and clamp is not None:
            continue   # linear + clamp: the plugin does not mask the gradient (saves no y); covered in test_bias_act_linear_clamp_grad
        assert_close(g[0], go[0], TOL[dtype], f'dx (create_graph={create_graph})')
        if has_b:
            assert_close(g[1], go[1], max(TOL[dtype], 2e-5), f'db (create_graph={create_graph})')
    # second order: d<dx, v>/d(dy) = grad-1 pass of v
    if act == 'lrelu':
        dyv = dy.clone().requires_grad_(True)
        dx, = torch.autograd.grad(y, x, dyv, create_graph=True)
        v = torch.randn_like(dx)
        d_dy, = torch.autograd.grad(dx, dyv, v)
        alpha, g_, c_ = R._resolve(act, None, gain, clamp)
        ref = R.bias_act_kernel(v.float().cpu(), None, None, y.detach().float().cpu(), None, 1, 1, act, alpha, g_, c_)
        assert_close(d_dy, ref, TOL[dtype], 'd_dy')


@pytest.mark.parametrize('dtype,cl', [(torch.float32, False), (torch.float16, True)])
def test_bias_act_fused_grad_is_differentiable_through_db(dtype, cl):
    """Under create_graph the backward is ONE fused (dx, db) pass (R1 never uses db, and `dx.sum` over the activation would cost a pass of
    its own); it stays differentiable: d<dx, v> + <db, u> / d(dy) = slope(y) * (v + u[c])."""
    torch.manual_seed(5)
    x = torch.randn(3, 16, 9, 7, device=DEV).to(dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    b = torch.randn(16, device=DEV).to(dtype).requires_grad_(True)
    y = bias_act.bias_act(x, b, act='lrelu', clamp=1.5)
    dy = torch.randn_like(y).requires_grad_(True)
    dx, db = torch.autograd.grad(y, [x, b], dy, create_graph=True)
    assert_close(db, dx.detach().float().sum([0, 2, 3]).cpu(), 1e-2 if dtype == torch.float16 else 2e-5, 'db = sum dx')
    v, u = torch.randn_like(dx), torch.randn_like(db)
    d_dy, = torch.autograd.grad([dx, db], dy, [v, u])
    yd = y.detach().float()
    slope = torch.where(yd > 0, torch.full_like(yd, float(np.sqrt(2))), torch.full_like(yd, float(np.sqrt(2)) * 0.2)) * (yd.abs() < 1.5)
    ref = slope * (v.float() + u.float().reshape(1, -1, 1, 1))
    assert_close(d_dy, ref.cpu(), 1e-2 if dtype == torch.float16 else 2e-5, 'd_dy')
    only_db, = torch.autograd.grad(torch.autograd.grad(y, [b], dy, create_graph=True)[0], dy, u)
    assert_close(only_db, (slope * u.float().reshape(1, -1, 1, 1)).cpu(), 1e-2 if dtype == torch.float16 else 2e-5, 'd_dy from db alone')


def test_bias_act_linear_clamp_grad():
    """Reference CUDA semantics: `linear` saves no output, so the clamp does not mask its gradient (OPS/bias_act.py:151-154)."""
    x = (torch.randn(2, 3, 4, 4, device=DEV) * 300).requires_grad_(True)
    y = bias_act.bias_act(x, None, act='linear', clamp=256.0)
    assert float(y.abs().max()) <= 256.0
    dy = torch.randn_like(y)
    dx, = torch.autograd.grad(y, x, dy)
    assert_close(dx, dy.cpu(), 1e-6, 'dx')


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64, torch.float16])
@pytest.mark.parametrize('act', gg.ALL_ACTS)
def test_bias_act_all_activations(golden, act, dtype):
    G = golden('ops_bias_act.npz')
    x = t(G[f'bias_act_all/{act}/x'], dtype).requires_grad_(True)
    b = t(G[f'bias_act_all/{act}/b'], dtype)
    dy = t(G[f'bias_act_all/{act}/dy'], dtype).requires_grad_(True)
    v = t(G[f'bias_act_all/{act}/v'], dtype)
    tol = {torch.float32: 2e-5, torch.float64: 2e-7, torch.float16: 1e-2}[dtype]   # fp64: `gain` crosses the C ABI as a float, as in OPS/bias_act.cpp:32
    y = bias_act.bias_act(x, b, dim=1, act=act)
    assert_close(y, G[f'bias_act_all/{act}/y'], tol, 'y')
    dx, = torch.autograd.grad(y, x, dy, create_graph=True)
    assert_close(dx, G[f'bias_act_all/{act}/dx'], tol, 'dx')
    d_x, d_dy = torch.autograd.grad(dx, [x, dy], v, allow_unused=True)
    assert_close(d_dy, G[f'bias_act_all/{act}/d_dy'], tol, 'd_dy')
    if R.ACTIVATIONS[act][4] and dtype != torch.float16:
        assert_close(d_x, G[f'bias_act_all/{act}/d_x'], tol * 10, 'd_x')


def test_bias_act_plugin_interface_and_errors():
    """The object returned by get_plugin has the reference's pybind signature (OPS/bias_act.cpp:32) and its checks."""
    plugin = custom_ops.get_plugin('bias_act_plugin', sources=['bias_act.cpp', 'bias_act.cu'], headers=['bias_act.h'], source_dir='.')
    x = torch.randn(4, 8, 5, 5, device=DEV)
    b = torch.randn(8, device=DEV)
    e = torch.empty([0])
    y = plugin.bias_act(x, b, e, e, e, 0, 1, 3, 0.2, 2 ** 0.5, 256.0)
    assert_close(y, R.bias_act(x.cpu(), b.cpu(), act='lrelu', clamp=256.0), 1e-6)
    with pytest.raises(RuntimeError):
        plugin.bias_act(x, torch.randn(7, device=DEV), e, e, e, 0, 1, 3, 0.2, 1.0, -1.0)       # wrong bias length
    with pytest.raises(RuntimeError):
        plugin.bias_act(x.cpu(), b.cpu(), e, e, e, 0, 1, 3, 0.2, 1.0, -1.0)                    # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        plugin.bias_act(x[:, :, ::2], b, e, e, e, 0, 1, 3, 0.2, 1.0, -1.0)                     # not dense
    with pytest.raises(RuntimeError):
        custom_ops.get_plugin('filtered_lrelu_plugin', sources=[])
    empty = torch.empty(0, 8, 5, 5, device=DEV)
    assert bias_act.bias_act(empty, b, act='lrelu').shape == empty.shape                       # empty input


@pytest.mark.parametrize('shape,cl', [((32, 64, 256, 256), True), ((32, 64, 256, 256), False), ((8, 512, 32, 32), True), ((3, 7, 33, 17), False)])
def test_bias_act_full_size_properties(shape, cl):
    """BASELINE.json sizes: compare with the same formula evaluated by torch elementwise ops on the device (an
    independent code path), plus db == dx.sum, in fp16."""
    torch.manual_seed(0)
    x = torch.randn(shape, device=DEV, dtype=torch.float16)
    b = torch.randn(shape[1], device=DEV, dtype=torch.float16)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    b.requires_grad_(True)
    gain, clamp = float(np.sqrt(2)), 2.0
    y = bias_act.bias_act(x, b, act='lrelu', gain=gain, clamp=clamp)
    u = x.detach().float() + b.detach().float().reshape(1, -1, 1, 1)
    ref = (torch.nn.functional.leaky_relu(u, 0.2) * gain).clamp(-clamp, clamp)
    assert_close(y, ref.half(), 2e-3, 'y')
    dy = torch.randn_like(y)
    dx, db = torch.autograd.grad(y, [x, b], dy)
    mask = (ref.abs() < clamp).float()
    refdx = dy.float() * gain * torch.where(u > 0, 1.0, 0.2) * mask
    # saturated elements sit exactly on the clamp after rounding; exclude the measure-zero boundary set
    inner = ((y.detach().float().abs() - clamp).abs() > 1e-2)
    assert_close(dx.float() * inner, refdx * inner, 2e-3, 'dx')
    assert_close(db.float(), dx.float().sum([0, 2, 3]), 5e-3, 'db')


@pytest.mark.parametrize('shape,cl,dtype', [((32, 64, 256, 256), True, torch.float16), ((32, 64, 128, 128), False, torch.float16),
                                            ((16, 128, 65, 33), True, torch.float16), ((8, 512, 32, 32), True, torch.float32),
                                            ((5, 64, 129, 131), True, torch.float16), ((1, 1, 1237, 1733), False, torch.float16)])
def test_bias_act_bulk_staged_equals_direct(shape, cl, dtype):
    """The bulk-copy staged streaming kernels (csrc/stream_bulk.cuh) and the direct vector kernels compute the same
    thing bit for bit: forward, dx and db, incl. sizes that are not a whole number of chunks / vectors."""
    from gan_track_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(1)
    x = torch.randn(shape, device=DEV).to(dtype)
    b = torch.randn(shape[1], device=DEV).to(dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    dy = torch.randn_like(x)
    out = []
    for variant in (0, 1):
        old = lib.gt_stream_config(variant)
        try:
            xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
            y = bias_act.bias_act(xr, br, act='lrelu', gain=float(np.sqrt(2)), clamp=1.5)
            dx, db = torch.autograd.grad(y, [xr, br], dy)
            y2 = bias_act.bias_act(xr.detach(), None, act='linear', gain=0.5)
            out.append((y.detach(), dx, db, y2))
        finally:
            lib.gt_stream_config(old)
    torch.cuda.synchronize()
    for a, c, name in zip(out[0], out[1], ['y', 'dx', 'db', 'y_linear']):
        if name == 'db':
            assert_close(a.float(), c.float(), 2e-3, name)       # different partial-sum grouping
        else:
            assert torch.equal(a, c), name


# ---------------------------------------------------------------------------------------------------- upfirdn2d

@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.float64])
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('case', gg.UPFIRDN_CASES, ids=lambda c: c[0])
def test_upfirdn2d_cases(golden, case, cl, dtype):
    name, f, kw, shape = case
    G = golden('ops_upfirdn2d.npz')
    x = t(G[f'upfirdn2d/{name}/x'], dtype)
    if cl:
        # widen channels to a multiple of 8 so the channels-last vector kernel is the one exercised
        reps = 8
        x = x.repeat(1, reps, 1, 1).contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    ft = upfirdn2d.setup_filter(f, device=DEV) if f is not None else None
    y = upfirdn2d.upfirdn2d(x, ft, **kw)
    yo = R.upfirdn2d(x.detach().float().cpu(), ft.cpu() if ft is not None else None, **kw)
    assert y.shape == yo.shape
    assert_close(y, yo, TOL[dtype] if dtype != torch.float64 else 1e-6, 'y')
    if dtype == torch.float32 and not cl:
        assert_close(y, G[f'upfirdn2d/{name}/y'], 1e-5, 'y vs golden')
    dy = torch.randn_like(y)
    dx, = torch.autograd.grad(y, x, dy)
    xo = x.detach().float().cpu().requires_grad_(True)
    dxo, = torch.autograd.grad(R.upfirdn2d(xo, ft.cpu() if ft is not None else None, **kw), xo, dy.float().cpu())
    assert_close(dx, dxo, TOL[dtype] if dtype != torch.float64 else 1e-6, 'dx')
    if cl:
        assert y.is_contiguous(memory_format=torch.channels_last)


def test_upfirdn2d_wrappers_strides_and_errors(golden):
    G = golden('ops_upfirdn2d.npz')
    x = t(G['upfirdn2d/wrappers/x'])
    f4 = upfirdn2d.setup_filter(gg.F4, device=DEV)
    assert_close(upfirdn2d.filter2d(x, f4), G['upfirdn2d/wrappers/filter2d'], 1e-5)
    assert_close(upfirdn2d.upsample2d(x, f4), G['upfirdn2d/wrappers/upsample2d'], 1e-5)
    assert_close(upfirdn2d.downsample2d(x[:, :, :, :6], f4), G['upfirdn2d/wrappers/downsample2d'], 1e-5)    # strided view input
    # arbitrary strides: transposed view
    xt = torch.randn(2, 3, 9, 7, device=DEV).transpose(2, 3)
    assert_close(upfirdn2d.upfirdn2d(xt, f4, padding=1), R.upfirdn2d(xt.cpu(), f4.cpu(), padding=1), 1e-5)
    # DC gain of upsample2d is 1 away from the border
    ones = torch.ones(1, 1, 16, 16, device=DEV)
    assert torch.allclose(upfirdn2d.upsample2d(ones, f4)[:, :, 3:-3, 3:-3], torch.ones(1, 1, 26, 26, device=DEV), atol=1e-6)
    sym6 = upfirdn2d.setup_filter(gg.SYM6, device=DEV)
    assert torch.allclose(upfirdn2d.upsample2d(ones, sym6)[:, :, 8:-8, 8:-8], torch.ones(1, 1, 16, 16, device=DEV), atol=1e-5)
    plugin = custom_ops.get_plugin('upfirdn2d_plugin', sources=['upfirdn2d.cpp', 'upfirdn2d.cu'], headers=['upfirdn2d.h'], source_dir='.')
    with pytest.raises(RuntimeError):
        plugin.upfirdn2d(x, f4.double(), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)         # f must be float32
    with pytest.raises(RuntimeError):
        plugin.upfirdn2d(x[:, :, :2, :2], f4, 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)    # output smaller than 1x1
    with pytest.raises(RuntimeError):
        plugin.upfirdn2d(x.cpu(), f4.cpu(), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)      # CPU tensor: no fallback


@pytest.mark.parametrize('N,C,H,cl', [(8, 64, 257, True), (8, 64, 257, False), (4, 512, 33, True)])
def test_upfirdn2d_full_size_vs_depthwise_conv(N, C, H, cl):
    """G conv0's blur at full size: [N,C,2H+1,2H+1] -> [N,C,2H,2H], 4x4 taps, pad 1, gain 4; fp16.  Checked against a
    depthwise torch convolution in fp32 (independent code path) and by linearity."""
    torch.manual_seed(1)
    x = torch.randn(N, C, H, H, device=DEV, dtype=torch.float16)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    f4 = upfirdn2d.setup_filter(gg.F4, device=DEV)
    y = upfirdn2d.upfirdn2d(x, f4, padding=[1, 1, 1, 1], gain=4)
    w = (f4.flip([0, 1]) * 4)[None, None].repeat(C, 1, 1, 1)
    ref = torch.nn.functional.conv2d(x.float(), w, padding=1, groups=C)
    assert y.shape == ref.shape
    assert_close(y, ref, 2e-3, 'y')
    y2 = upfirdn2d.upfirdn2d(x * 0.5, f4, padding=[1, 1, 1, 1], gain=4)
    assert_close(y2.float() * 2, y.float(), 2e-3, 'linearity')


STRIDED_TMA_CASES = [
    # (up, down, padding, N, C, H, W)
    (2, 1, [2, 1, 2, 1], 2, 64, 96, 96),      # backward of the skip downsample / upsample2d
    (2, 1, [1, 2, 1, 2], 2, 64, 67, 45),      # odd parity, ragged tiles
    (2, 1, [3, 0, 2, 1], 1, 128, 50, 41),     # mixed parity
    (2, 1, [0, 0, 0, 0], 2, 64, 64, 80),
    (2, 1, [-1, 2, 1, -2], 2, 64, 72, 64),    # cropping
    (1, 2, [1, 1, 1, 1], 2, 64, 128, 128),    # the skip downsample (OPS/conv2d_resample.py:94-97)
    (1, 2, [2, 2, 2, 2], 2, 64, 131, 97),
    (1, 2, [0, 1, 3, 0], 1, 128, 100, 90),
    (1, 2, [-1, 1, 2, -1], 2, 64, 96, 130),
]


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
@pytest.mark.parametrize('flip', [False, True])
@pytest.mark.parametrize('case', STRIDED_TMA_CASES, ids=lambda c: f'up{c[0]}down{c[1]}pad{"_".join(map(str, c[2]))}')
def test_upfirdn2d_strided_tma_vs_oracle(case, flip, dtype):
    """Factor-2 resampling on channels-last tensors large enough for the TMA-staged kernels (upfirdn2d_tma_strided.cu),
    with an asymmetric 4x4 filter so that tap orientation, flip and the phase/tap pairing are all pinned by the oracle."""
    up, down, padding, N, C, H, W = case
    if dtype == torch.float32:
        C //= 2 if C > 64 else 1
    torch.manual_seed(3)
    x = torch.randn(N, C, H, W, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    f = torch.randn(4, 4, device=DEV)
    kw = dict(up=up, down=down, padding=padding, flip_filter=flip, gain=1.7)
    y = upfirdn2d.upfirdn2d(x, f, **kw)
    yo = R.upfirdn2d(x.detach().float().cpu(), f.cpu(), **kw)
    assert y.shape == yo.shape
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert_close(y, yo, TOL[dtype], 'y')
    dy = torch.randn_like(y)
    dx, = torch.autograd.grad(y, x, dy)          # the other strided kernel
    xo = x.detach().float().cpu().requires_grad_(True)
    dxo, = torch.autograd.grad(R.upfirdn2d(xo, f.cpu(), **kw), xo, dy.float().cpu())
    assert_close(dx, dxo, TOL[dtype], 'dx')


@pytest.mark.parametrize('up,down,shape', [(1, 2, (32, 64, 256, 256)), (2, 1, (32, 64, 128, 128)), (1, 2, (32, 512, 32, 32)), (2, 1, (16, 256, 32, 32))])
def test_upfirdn2d_strided_tma_equals_direct_full_size(up, down, shape):
    """BASELINE.json sizes: TMA-staged vs the thread-per-output kernel (an independent code path) + adjointness
    <upfirdn(x), dy> == <x, upfirdn_backward(dy)>."""
    from gan_track_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(4)
    x = torch.randn(shape, device=DEV, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    f4 = upfirdn2d.setup_filter(gg.F4, device=DEV)
    kw = dict(up=up, down=down, padding=[2, 1, 2, 1] if up == 2 else [1, 1, 1, 1], gain=up * up)
    out = []
    for variant in (0, 1):
        old = lib.gt_stream_config(variant)
        try:
            out.append(upfirdn2d.upfirdn2d(x, f4, **kw))
        finally:
            lib.gt_stream_config(old)
    assert_close(out[0], out[1], 1e-3, 'tma vs direct')
    xr = x.clone().requires_grad_(True)
    y = upfirdn2d.upfirdn2d(xr, f4, **kw)
    dy = torch.randn_like(y)
    dx, = torch.autograd.grad(y, xr, dy)
    lhs = (y.detach().double() * dy.double()).sum()
    rhs = (x.double() * dx.double()).sum()
    assert abs(lhs - rhs) <= 1e-2 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)


# ---------------------------------------------------------------------------------------------------- conv family

@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('case', gg.CONV_RESAMPLE_CASES, ids=lambda c: c[0])
def test_conv2d_resample_cases(golden, case, dtype):
    name, ci, co, k, kw, h = case
    G = golden('ops_conv.npz')
    x = t(G[f'conv2d_resample/{name}/x'], dtype).requires_grad_(True)
    w = t(G[f'conv2d_resample/{name}/w'], dtype).requires_grad_(True)
    dy = t(G[f'conv2d_resample/{name}/dy'], dtype)
    f4 = upfirdn2d.setup_filter(gg.F4, device=DEV)
    y = conv2d_resample.conv2d_resample(x, w, f=f4, **kw)
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert_close(y, G[f'conv2d_resample/{name}/y'], tol, 'y')
    dx, dw = torch.autograd.grad(y, [x, w], dy)
    assert_close(dx, G[f'conv2d_resample/{name}/dx'], tol, 'dx')
    assert_close(dw, G[f'conv2d_resample/{name}/dw'], tol, 'dw')


@pytest.mark.parametrize('transpose,stride,pad', [(False, 1, 1), (False, 2, 0), (True, 2, 0), (False, 1, 0)])
def test_conv_gradfix_double_backward(transpose, stride, pad):
    """Second-order gradients of the conv Function family against torch's native double backward (fp64-free: fp32)."""
    torch.manual_seed(2)
    ci, co, k = 6, 5, 3
    x = torch.randn(2, ci, 9, 9, device=DEV, requires_grad=True)
    w = torch.randn((ci, co, k, k) if transpose else (co, ci, k, k), device=DEV, requires_grad=True)

    def run(fn_conv, fn_convT):
        y = (fn_convT if transpose else fn_conv)(x, w, stride=stride, padding=pad)
        gy = torch.randn(y.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
        gx, gw = torch.autograd.grad(y, [x, w], gy, create_graph=True)
        s = gx.square().sum() + (gw * gw.detach().sign()).sum()
        ggx, ggw = torch.autograd.grad(s, [x, w])
        return y, gx, gw, ggx, ggw

    ours = run(conv2d_gradfix.conv2d, conv2d_gradfix.conv_transpose2d)
    theirs = run(torch.nn.functional.conv2d, torch.nn.functional.conv_transpose2d)
    for a, b, name in zip(ours, theirs, ['y', 'gx', 'gw', 'ggx', 'ggw']):
        assert_close(a, b, 2e-5, name)
    # no_weight_gradients() suppresses dw in first-order backward
    y = conv2d_gradfix.conv2d(x, w if not transpose else w.transpose(0, 1).contiguous(), padding=1)
    with conv2d_gradfix.no_weight_gradients():
        gx, gw = torch.autograd.grad(y.sum(), [x, w], allow_unused=True)
    assert gx is not None and gw is None


def test_fma_and_grid_sample(golden):
    G = golden('ops_conv.npz')
    a, b, c = (t(G['fma/a']).requires_grad_(True), t(G['fma/b']).requires_grad_(True), t(G['fma/c']).requires_grad_(True))
    y = fma.fma(a, b, c)
    assert_close(y, G['fma/y'], 1e-6)
    ga, gb, gc = torch.autograd.grad(y.sum(), [a, b, c])
    assert ga.shape == a.shape and gb.shape == b.shape and gc.shape == c.shape
    assert_close(gb, a.detach().sum([2, 3], keepdim=True), 1e-5)
    img, grid = t(G['grid_sample/img']).requires_grad_(True), t(G['grid_sample/grid'])
    out = grid_sample_gradfix.grid_sample(img, grid)
    assert_close(out, G['grid_sample/y'], 1e-5)
    g1, = torch.autograd.grad(out.square().sum(), img, create_graph=True)
    g2, = torch.autograd.grad(g1.square().sum(), img)          # double backward must exist (R1 through the ADA pipe)
    assert torch.isfinite(g2).all()


# ---------------------------------------------------------------------------------------------------- fully-connected

@pytest.mark.parametrize('M,I,O,bias', [(32, 512, 512, True), (32, 8192, 512, True), (16, 512, 64, True), (4, 32, 32, False), (64, 516, 130, True),
                                        (7, 512, 1, True), (33, 64, 512, False)])
def test_fc_linear_matches_addmm_to_second_order(M, I, O, bias):
    """csrc/fc.cu against the reference's op sequence (S3/training/networks_stylegan2.py:115-126): value, first-order gradients of
    x / weight / bias, and the second-order terms the R1 and path-length passes need (gradient of <dx, v> w.r.t. dy-side inputs)."""
    from gan_track_b200.torch_utils.ops import fc
    torch.manual_seed(M * 7 + O)
    wg, bg = 1.0 / np.sqrt(I), 0.7
    x0 = torch.randn(M, I, device=DEV)
    w0 = torch.randn(O, I, device=DEV)
    b0 = torch.randn(O, device=DEV) if bias else None
    dy = torch.randn(M, O, device=DEV)
    v = torch.randn(M, I, device=DEV)
    outs = []
    for ours in (True, False):
        x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True) if bias else None
        assert fc.applicable(x, w)
        if ours:
            y = fc.linear(x, w, b, wg, bg)
        else:
            y = x.matmul((w * wg).t())
            if bias:
                y = y + (b * bg).unsqueeze(0)
        ins = [x, w] + ([b] if bias else [])
        g = torch.autograd.grad(y, ins, dy, create_graph=True)
        q = (g[0] * v).sum() + (g[1] * g[1]).sum() * 0.1
        g2 = torch.autograd.grad(q, [x, w], allow_unused=True)
        outs.append([y.detach()] + [t_.detach() for t_ in g] + [t_.detach() for t_ in g2 if t_ is not None])
    assert len(outs[0]) == len(outs[1])
    for a, c in zip(outs[0], outs[1]):
        assert a.shape == c.shape
        assert_close(a, c.cpu(), 2e-5, 'fc')


# ---------------------------------------------------------------------------------------------------- flat Adam / EMA

def test_flat_adam_and_ema_match_torch():
    """csrc/optim.cu against torch.optim.Adam (+ /num_gpus and nan_to_num of the gradient exchange) over several steps in which
    some parameters receive no gradient (they must be skipped: no moment decay, no step count), and torch._foreach_lerp_."""
    import copy
    from gan_track_b200.training import flat_optim
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(37, 129), torch.nn.Conv2d(3, 5, 3), torch.nn.Linear(11, 7, bias=False)).to(DEV)
    ref = copy.deepcopy(net)
    ema, ema_ref = copy.deepcopy(net), copy.deepcopy(net)
    kw = dict(lr=0.0025 * 0.8, betas=[0.0, 0.99 ** 0.8], eps=1e-8)
    opt_ref = torch.optim.Adam(ref.parameters(), **kw)
    fp, fe = flat_optim.FlatParams(net), flat_optim.FlatParams(ema)
    opt = flat_optim.FlatAdam(fp, **kw)
    params, rparams = list(net.parameters()), list(ref.parameters())
    for k, v in net.state_dict().items():
        assert torch.equal(v, ref.state_dict()[k])                 # re-homing kept the values
    for it in range(6):
        active = [0, 1, 2, 3, 4] if it % 3 else [0, 2, 4]          # a "reg" phase touches fewer parameters
        grads = [torch.randn_like(rparams[i]) * (10.0 ** (it % 3 - 1)) for i in active]
        grads[0][0, 0] = float('nan')
        grads[1].view(-1)[1] = float('inf')
        for p in rparams:
            p.grad = None
        for i, g in zip(active, grads):
            rparams[i].grad = torch.nan_to_num(g / 2, nan=0, posinf=1e5, neginf=-1e5)
        opt_ref.step()
        flat = torch.cat([g.flatten() for g in grads])
        opt.step('main' if it % 3 else 'reg', active, flat, grad_scale=0.5)
        with torch.no_grad():
            torch._foreach_lerp_(list(ema_ref.parameters()), rparams, 0.03)
        flat_optim.ema_update(fe, fp, 0.03)
    for p, r in zip(params, rparams):
        assert_close(p, r.detach().cpu(), 2e-6, 'adam param')
    for p, r in zip(ema.parameters(), ema_ref.parameters()):
        assert_close(p, r.detach().cpu(), 2e-6, 'ema param')


# ---------------------------------------------------------------------------------------------------- FromRGB (1 channel)

@pytest.mark.parametrize('dtype,C,act,clamp', [(torch.float16, 64, 'lrelu', 256.0), (torch.float16, 128, 'lrelu', 0.5), (torch.float32, 16, 'lrelu', None),
                                               (torch.float16, 8, 'linear', 1.0)])
def test_fromrgb1_matches_conv_bias_act(dtype, C, act, clamp):
    """csrc/rgb.cu against the reference's op sequence conv2d (1x1, one input channel) + bias_act: value, first-order gradients
    (fused pass) and the R1-style second order (tensor-op form under create_graph)."""
    from gan_track_b200.torch_utils.ops import rgb
    torch.manual_seed(C)
    N, H, W = 3, 37, 41
    x0 = torch.randn(N, 1, H, W, device=DEV).to(dtype)
    w0 = (torch.randn(C, 1, 1, 1, device=DEV) * 1.3).to(dtype)
    b0 = torch.randn(C, device=DEV).to(dtype)
    dy = torch.randn(N, C, H, W, device=DEV).to(dtype)
    tol = 2e-3 if dtype == torch.float16 else 2e-5
    outs = []
    for ours in (True, False):
        x, w, b = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        if ours:
            assert rgb.applicable(x, C)
            y = rgb.fromrgb1(x, w.reshape(-1), b, act=act, clamp=clamp)
            assert y.stride(1) == 1 or C == 1
        else:
            y = bias_act.bias_act(torch.nn.functional.conv2d(x, w), b, act=act, clamp=clamp)
        g = torch.autograd.grad(y, [x, w, b], dy, retain_graph=True)                       # fused first-order pass
        gx, = torch.autograd.grad(y, [x], dy, create_graph=True)                           # R1: gradient w.r.t. the image ...
        g2 = torch.autograd.grad(gx.float().square().sum(), [w], allow_unused=True)        # ... differentiated w.r.t. the weights
        outs.append([y.detach()] + [t_.detach().reshape(-1) for t_ in g] + [gx.detach().reshape(-1)] + [t_.detach().reshape(-1) for t_ in g2 if t_ is not None])
    assert len(outs[0]) == len(outs[1])
    for a, c in zip(outs[0], outs[1]):
        assert a.shape == c.shape
        assert_close(a.float(), c.float().cpu(), tol if a.numel() > C else 10 * tol, 'fromrgb1')


@pytest.mark.parametrize('dtype,C,clamp', [(torch.float16, 64, 256.0), (torch.float16, 128, 0.3), (torch.float32, 32, None), (torch.float16, 256, 256.0)])
def test_torgb1_matches_modulated_conv_bias_act(dtype, C, clamp):
    """csrc/rgb.cu (ToRGB with one image channel) against the op-by-op form (modulate, 1x1 convolution, bias_act): value, the
    first-order gradients of x / weight / styles / bias (fused pass) and a path-length-style second order (under create_graph)."""
    from gan_track_b200.torch_utils.ops import rgb
    torch.manual_seed(C)
    N, H, W = 3, 24, 40
    x0 = torch.randn(N, C, H, W, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    w0 = torch.randn(1, C, 1, 1, device=DEV) / np.sqrt(C)
    s0 = torch.randn(N, C, device=DEV) * 0.5 + 1
    b0 = torch.randn(1, device=DEV) * 0.1
    dy = torch.randn(N, 1, H, W, device=DEV).to(dtype)
    tol = 4e-3 if dtype == torch.float16 else 2e-5
    outs = []
    for ours in (True, False):
        x, w, s, b = (t_.clone().requires_grad_(True) for t_ in (x0, w0, s0, b0))
        if ours:
            assert rgb.torgb_applicable(x, 1)
            y = rgb.torgb1(x, w, s, b, clamp=clamp)
        else:
            y = rgb._torgb_reference(x, w, s, b, clamp)
        g = torch.autograd.grad(y, [x, w, s, b], dy, retain_graph=True)
        gs, = torch.autograd.grad(y, [s], dy, create_graph=True)                   # path length: gradient w.r.t. the styles ...
        g2 = torch.autograd.grad(gs.square().sum(), [w, x], allow_unused=True)      # ... differentiated w.r.t. weight and x
        outs.append([y.detach()] + [t_.detach() for t_ in g] + [gs.detach()] + [t_.detach() for t_ in g2 if t_ is not None])
    assert len(outs[0]) == len(outs[1])
    for i, (a, c) in enumerate(zip(outs[0], outs[1])):
        assert a.shape == c.shape
        assert_close(a.float(), c.float().cpu(), tol if i != 4 else 10 * tol, f'torgb1 output {i}')


def test_torgb_op_by_op_matmul_route_matches_convolution_route():
    """The path-length pass keeps ToRGB in op-by-op form; on channels-last tensors its 1x1 convolution to one channel runs as a
    per-pixel matmul on a free view.  Same values / first- and second-order gradients as the convolution route (NCHW input)."""
    from gan_track_b200.torch_utils.ops import rgb
    from gan_track_b200.training import networks_stylegan2 as nets
    torch.manual_seed(11)
    layer = nets.ToRGBLayer(64, 1, w_dim=32, conv_clamp=256, channels_last=True).to(DEV)
    x0 = torch.randn(3, 64, 24, 20, device=DEV).half()
    w0 = torch.randn(3, 32, device=DEV)
    probe = torch.randn(3, 1, 24, 20, device=DEV)
    res = []
    for cl in (True, False):
        x = (x0.contiguous(memory_format=torch.channels_last) if cl else x0.contiguous()).requires_grad_(True)
        w = w0.clone().requires_grad_(True)
        with rgb.op_by_op_torgb():
            y = layer(x, w, fused_modconv=False)
            gw, = torch.autograd.grad((y.float() * probe).sum(), [w], create_graph=True)
            pen = gw.square().sum()
            g2 = torch.autograd.grad(pen, [x, w, layer.weight, layer.affine.weight], allow_unused=True)
        # y is linear in the styles, so d(pen)/dw vanishes: autograd reports it as None on one route and as zeros on the other
        assert g2[1] is None or float(g2[1].abs().max()) == 0.0
        res.append((y.detach(), gw.detach(), g2[0], g2[2], g2[3]))
    for a, b in zip(res[0], res[1]):
        assert_close(a, b, 1e-2)
