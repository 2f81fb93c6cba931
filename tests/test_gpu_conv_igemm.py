"""GPU: the tcgen05 implicit-GEMM convolutions (csrc/conv_igemm.cu, csrc/conv_wgrad.cu) through the C ABI against the
arithmetic the reference runs -- torch's own convolution (OPS/conv2d_gradfix.py:40,45) -- evaluated in fp32 on the same
fp16 inputs.  Tolerance: fp16 layers within 1e-2 relative (north star); observed <= 5e-4."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-3   # well inside the 1e-2 the north star allows for fp16 layers

# (name, N, Cin, Cout, H, W, k, stride, pad, transpose) -- every conv flavour on the StyleGAN2 path and
# its data gradient, ragged sizes (33, 257 = 2H+1 of the up path), tiny maps (4x4, 8x8: several images per tile).
CASES = [
    ('3x3_p1_64_64_32', 2, 64, 64, 32, 32, 3, 1, 1, False),
    ('3x3_p1_128_256_16', 4, 128, 256, 16, 16, 3, 1, 1, False),
    ('3x3_p1_64_128_33', 3, 64, 128, 33, 33, 3, 1, 1, False),
    ('3x3_p1_64_64_8', 8, 64, 64, 8, 8, 3, 1, 1, False),
    ('3x3_p1_64_64_4', 5, 64, 64, 4, 4, 3, 1, 1, False),
    ('3x3_p1_512_512_32', 2, 512, 512, 32, 32, 3, 1, 1, False),
    ('1x1_128_64_32', 2, 128, 64, 32, 32, 1, 1, 0, False),
    ('3x3_s2_64_128_33', 2, 64, 128, 33, 33, 3, 2, 0, False),
    ('3x3_s2_64_64_257', 1, 64, 64, 257, 257, 3, 2, 0, False),
    ('3x3_T_s2_128_64_16', 2, 128, 64, 16, 16, 3, 2, 0, True),
    ('3x3_T_s2_64_64_128', 1, 64, 64, 128, 128, 3, 2, 0, True),
    ('3x3_T_s1_p1_64_128_32', 2, 64, 128, 32, 32, 3, 1, 1, True),
    ('3x3_T_s1_p0_64_64_16', 2, 64, 64, 16, 16, 3, 1, 0, True),
    # stride-1 3x3 with enough pixels for the halo-staged tap-paired wgrad kernel (csrc/conv_wgrad_halo.cu), incl. ragged tiles
    ('3x3_p1_64_64_128_n8', 8, 64, 64, 128, 128, 3, 1, 1, False),
    ('3x3_p1_128_64_72x64', 4, 128, 64, 72, 64, 3, 1, 1, False),
    ('3x3_T_s1_p1_64_128_40', 6, 64, 128, 40, 40, 3, 1, 1, True),
    ('3x3_p0_64_64_66', 4, 64, 64, 66, 66, 3, 1, 0, False),
    ('3x3_p1_256_256_32_n8', 8, 256, 256, 32, 32, 3, 1, 1, False),
    # stride-2 / transposed stride-2 with enough pixels for the parity-plane staging of the halo wgrad kernel (>= 64 tiles of 8x8 output
    # pixels), incl. even input sizes (2 OH + 2 rows), widths that are no multiple of 8 and tile rows past the image
    ('3x3_s2_64_128_129_n4', 4, 64, 128, 129, 129, 3, 2, 0, False),
    ('3x3_s2_128_64_66x90_n5', 5, 128, 64, 66, 90, 3, 2, 0, False),
    ('3x3_T_s2_128_64_36x44_n3', 3, 128, 64, 36, 44, 3, 2, 0, True),
    ('3x3_T_s2_256_128_32_n4', 4, 256, 128, 32, 32, 3, 2, 0, True),
    # stride-1 3x3 with >= 128 U channels and enough pixel tiles per CTA for the wide (N = 128, two CTA types) weight-gradient kernel
    # (csrc/conv_wgrad_halo_wide.cu), incl. ragged tiles and the transposed form (U = x)
    ('3x3_p1_128_128_96_n8', 8, 128, 128, 96, 96, 3, 1, 1, False),
    ('3x3_p1_256_128_40x72_n8', 8, 256, 128, 40, 72, 3, 1, 1, False),
    ('3x3_T_s1_p1_128_256_56_n8', 8, 128, 256, 56, 56, 3, 1, 1, True),
    # ... and its stride-2 (parity-plane) form
    ('3x3_s2_256_512_33_n24', 24, 256, 512, 33, 33, 3, 2, 0, False),
    ('3x3_T_s2_512_256_16_n24', 24, 512, 256, 16, 16, 3, 2, 0, True),
]


def _inputs(N, ci, co, H, W, k, tr, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn([N, ci, H, W], device='cuda', generator=g).to(torch.float16).contiguous(memory_format=torch.channels_last)
    wshape = [ci, co, k, k] if tr else [co, ci, k, k]
    w = (torch.randn(wshape, device='cuda', generator=g) / (ci * k * k) ** 0.5).to(torch.float16)
    return x, w


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_igemm_forward_matches_torch(case):
    from gan_track_b200.torch_utils.ops import conv_igemm
    _, N, ci, co, H, W, k, s, p, tr = case
    x, w = _inputs(N, ci, co, H, W, k, tr)
    y = conv_igemm.igemm_forward(x, w, transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
    assert y is not None, 'case must be covered by the tcgen05 kernel'
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv_transpose2d(x.float(), w.float(), stride=s, padding=p) if tr else F.conv2d(x.float(), w.float(), stride=s, padding=p)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert y.shape == ref.shape and y.dtype == torch.float16
    assert _rel(y, ref) <= TOL


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_igemm_wgrad_matches_torch(case):
    from gan_track_b200.torch_utils.ops import conv_igemm
    _, N, ci, co, H, W, k, s, p, tr = case
    x, w = _inputs(N, ci, co, H, W, k, tr, seed=1)
    OH, OW = conv_igemm.out_size(H, W, k, k, s, p, tr)
    dy = torch.randn([N, co, OH, OW], device='cuda').to(torch.float16).contiguous(memory_format=torch.channels_last)
    dw = conv_igemm.igemm_wgrad(dy, x, tuple(w.shape), transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
    assert dw is not None
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        _, ref, _ = torch.ops.aten.convolution_backward(dy.float(), x.float(), w.float(), None, [s, s], [p, p], [1, 1], tr, [0, 0], 1,
                                                        [False, True, False])
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert dw.shape == ref.shape
    assert _rel(dw, ref) <= TOL
    # deterministic: the split-K reduction has a fixed order
    dw2 = conv_igemm.igemm_wgrad(dy, x, tuple(w.shape), transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
    assert torch.equal(dw, dw2)
    # the per-tap-row kernel and the halo-staged tap-paired kernel agree (same products, different summation split)
    from gan_track_b200 import _lib
    old_v = _lib.load().gt_conv_wgrad_config(1)
    try:
        dw3 = conv_igemm.igemm_wgrad(dy, x, tuple(w.shape), transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
    finally:
        _lib.load().gt_conv_wgrad_config(old_v)
    assert _rel(dw3, ref) <= TOL and _rel(dw, dw3) <= TOL


def test_igemm_linearity_and_adjointness_full_size():
    """Size-independent properties at a training shape: conv is linear in x, and <conv(x), dy> == <x, dgrad(dy)> == <w, wgrad>."""
    from gan_track_b200.torch_utils.ops import conv_igemm
    N, ci, co, H = 8, 128, 128, 128
    x, w = _inputs(N, ci, co, H, H, 3, False, seed=2)
    kw = dict(output_padding=(0, 0), stride=(1, 1), padding=(1, 1), groups=1)
    y = conv_igemm.igemm_forward(x, w, transpose=False, **kw).float()
    y2 = conv_igemm.igemm_forward(x * 0.5, w, transpose=False, **kw).float()          # power-of-two scaling is exact in fp16
    assert _rel(y2 * 2, y) <= 1e-3
    dy = torch.randn_like(y).to(torch.float16).contiguous(memory_format=torch.channels_last)
    dx = conv_igemm.igemm_forward(dy, w, transpose=True, **kw).float()
    dw = conv_igemm.igemm_wgrad(dy, x, tuple(w.shape), transpose=False, **kw).float()
    a = float((y.double() * dy.double()).sum())
    b = float((x.double() * dx.double()).sum())
    c = float((w.double() * dw.double()).sum())
    scale = float(y.double().norm() * dy.double().norm())
    assert abs(a - b) <= 2e-3 * scale and abs(a - c) <= 2e-3 * scale


def test_conv_backend_routes_fp16_path_shapes_to_igemm():
    from gan_track_b200.torch_utils.ops import conv2d_gradfix, conv_backend
    x, w = _inputs(2, 64, 64, 32, 32, 3, False)
    before = dict(conv_backend.stats)
    x.requires_grad_(True)
    w.requires_grad_(True)
    y = conv2d_gradfix.conv2d(x, w, padding=1)
    y.sum().backward()
    assert conv_backend.stats['igemm'] - before['igemm'] == 2            # forward + dgrad
    assert conv_backend.stats['igemm_wgrad'] - before['igemm_wgrad'] == 1
    assert conv_backend.stats['library'] == before['library']


# fp32 layers on the tensor cores as fp16 x 3 (csrc/conv_f16x3.cu + the fp32-output mode of the implicit-GEMM kernels)
FP32_CASES = [
    ('fp32_3x3_p1_512_512_4', 8, 512, 512, 4, 4, 3, 1, 1, False),
    ('fp32_3x3_p1_512_512_16', 4, 512, 512, 16, 16, 3, 1, 1, False),
    ('fp32_3x3_p1_513_512_4', 8, 513, 512, 4, 4, 3, 1, 1, False),           # the convolution after the minibatch-std layer (D b4)
    ('fp32_3x3_p1_64_128_19', 3, 64, 128, 19, 19, 3, 1, 1, False),
    ('fp32_3x3_T_s2_512_512_8', 4, 512, 512, 8, 8, 3, 2, 0, True),
    ('fp32_3x3_s2_512_512_17', 4, 512, 512, 17, 17, 3, 2, 0, False),
    ('fp32_3x3_s2_128_64_17', 3, 128, 64, 17, 17, 3, 2, 0, False),
    ('fp32_1x1_512_512_8', 5, 512, 512, 8, 8, 1, 1, 0, False),
    ('fp32_3x3_T_s1_p1_32_64_16', 2, 32, 64, 16, 16, 3, 1, 1, True),
    ('fp32_3x3_p1_40_72_9', 2, 40, 72, 9, 9, 3, 1, 1, False),               # ragged channel counts (padded to 64 / 128 inside)
]


@pytest.mark.parametrize('scale', [1.0, 1e-6, 3e3], ids=['unit', 'tiny', 'large'])
@pytest.mark.parametrize('case', FP32_CASES, ids=lambda c: c[0])
def test_fp32_conv_f16x3_accuracy(case, scale):
    """fp32 layers on the tensor cores (fp16 x 3 with power-of-two scaling, K-split partial sums) against a float64 reference: forward,
    data gradient and weight gradient within the north star's 1e-5 (fp32), for unit-scale tensors, for tiny gradients (1e-6: the
    second-order passes) and for large activations; no call may reach the library."""
    from gan_track_b200.torch_utils.ops import conv2d_gradfix, conv_backend
    _, N, ci, co, H, W, k, s, p, tr = case
    g = torch.Generator(device='cuda').manual_seed(5)
    x = (torch.randn([N, ci, H, W], device='cuda', generator=g) * scale).requires_grad_(True)
    wshape = [ci, co, k, k] if tr else [co, ci, k, k]
    w = (torch.randn(wshape, device='cuda', generator=g) / (ci * k * k) ** 0.5).requires_grad_(True)
    fn = conv2d_gradfix.conv_transpose2d if tr else conv2d_gradfix.conv2d
    before = dict(conv_backend.stats)
    y = fn(x, w, stride=s, padding=p)
    dy = torch.randn(y.shape, device='cuda', generator=g) * scale
    dx, dw = torch.autograd.grad(y, [x, w], dy)
    used = {k_: conv_backend.stats[k_] - before[k_] for k_ in before}
    assert used['library'] == 0 and used['library_wgrad'] == 0 and used['igemm'] == 2 and used['igemm_wgrad'] == 1, used
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    ref = F.conv_transpose2d(xr, wr, stride=s, padding=p) if tr else F.conv2d(xr, wr, stride=s, padding=p)
    rdx, rdw = torch.autograd.grad(ref, [xr, wr], dy.double())
    assert y.dtype == torch.float32 and y.shape == ref.shape and dw.shape == w.shape and dw.dtype == torch.float32

    def rel64(a, b):
        return float((a.detach().double() - b.detach()).abs().max() / b.detach().abs().max())
    errs = (rel64(y, ref), rel64(dx, rdx), rel64(dw, rdw))
    print(f'\n  f16x3 {case[0]} scale {scale:g}: y {errs[0]:.1e} dx {errs[1]:.1e} dw {errs[2]:.1e}')
    assert max(errs) <= 1e-5, errs


@pytest.mark.parametrize('act,gain,clamp,bias', [('lrelu', None, 256.0, True), ('lrelu', 0.7, 0.5, True), ('linear', 0.7071, 181.0, True), ('linear', None, None, False)])
@pytest.mark.parametrize('shape', [(4, 64, 64, 40, 40, 3, 1, 1), (3, 128, 256, 33, 33, 3, 2, 0), (2, 128, 64, 16, 16, 1, 1, 0),
                                   (2, 64, 64, 70, 200, 3, 1, 1), (1, 64, 64, 129, 257, 3, 1, 1)])        # row-streaming kernel, eight epilogue warps
def test_conv_bias_act_fused_epilogue_equals_unfused(shape, act, gain, clamp, bias):
    """bias_act fused into the convolution's epilogue (csrc/conv_common.cuh: conv_store32) against convolution followed by the
    bias_act kernel: identical values (same rounding points), identical first-order gradients, and the R1-style second order."""
    from gan_track_b200.torch_utils.ops import conv2d_gradfix
    N, ci, co, H, W, k, s, p = shape
    g = torch.Generator(device='cuda').manual_seed(11)
    x0 = torch.randn([N, ci, H, W], device='cuda', generator=g).to(torch.float16).contiguous(memory_format=torch.channels_last)
    w0 = (torch.randn([co, ci, k, k], device='cuda', generator=g) / (ci * k * k) ** 0.5).to(torch.float16)
    b0 = torch.randn([co], device='cuda', generator=g).to(torch.float16) if bias else None
    outs = []
    for fused in (True, False):
        old = conv2d_gradfix.fuse_bias_act
        conv2d_gradfix.fuse_bias_act = fused
        try:
            x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
            b = b0.clone().requires_grad_(True) if bias else None
            y = conv2d_gradfix.conv2d_bias_act(x, w, b, act=act, gain=gain, clamp=clamp, stride=s, padding=p)
            dy = torch.randn(y.shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(12)).to(torch.float16)
            ins = [x, w] + ([b] if bias else [])
            grads = torch.autograd.grad(y, ins, dy, create_graph=True)
            q = grads[0].float().square().sum()                           # R1: penalty on the input gradient ...
            with conv2d_gradfix.no_weight_gradients(False):
                g2 = torch.autograd.grad(q, [w], allow_unused=True)       # ... differentiated w.r.t. the weights
            outs.append([y.detach()] + [t.detach() for t in grads] + [t.detach() for t in g2 if t is not None])
        finally:
            conv2d_gradfix.fuse_bias_act = old
    assert len(outs[0]) == len(outs[1])
    assert torch.equal(outs[0][0], outs[1][0]), 'forward values must be bit-identical'
    for a, c in zip(outs[0][1:], outs[1][1:]):
        assert a.shape == c.shape
        assert _rel(a, c) <= 2e-3


@pytest.mark.parametrize('shape', [(2, 128, 64, 16, 16, 1, 1, 0), (3, 64, 128, 33, 33, 1, 1, 0), (1, 64, 64, 40, 136, 3, 1, 1), (2, 128, 256, 32, 32, 3, 1, 1)])
def test_conv_bias_act_fused_residual_equals_unfused(shape):
    """conv -> bias_act(linear, gain, clamp) -> + residual (DiscriminatorBlock's shortcut.add_(x), S3/training/networks_stylegan2.py:636) with the
    residual added in the convolution's epilogue: bit-identical to the three separate passes; gradients of x, w and the residual."""
    from gan_track_b200.torch_utils.ops import conv2d_gradfix
    N, ci, co, H, W, k, s, p = shape
    g = torch.Generator(device='cuda').manual_seed(21)
    x0 = torch.randn([N, ci, H, W], device='cuda', generator=g).to(torch.float16).contiguous(memory_format=torch.channels_last)
    w0 = (torch.randn([co, ci, k, k], device='cuda', generator=g) / (ci * k * k) ** 0.5).to(torch.float16)
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    a0 = (torch.randn([N, co, OH, OW], device='cuda', generator=g) * 2).to(torch.float16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn([N, co, OH, OW], device='cuda', generator=g).to(torch.float16).contiguous(memory_format=torch.channels_last)
    outs = []
    for fused in (True, False):
        old = conv2d_gradfix.fuse_bias_act
        conv2d_gradfix.fuse_bias_act = fused
        try:
            x, w, a = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), a0.clone().requires_grad_(True)
            before = dict(conv_backend_stats())
            y = conv2d_gradfix.conv2d_bias_act(x, w, None, act='linear', gain=0.70710678, clamp=1.5, stride=s, padding=p, addend=a * 1.0)
            outs.append([y.detach()] + [t.detach() for t in torch.autograd.grad(y, [x, w, a], dy)])
        finally:
            conv2d_gradfix.fuse_bias_act = old
    assert torch.equal(outs[0][0], outs[1][0]), 'forward values must be bit-identical'
    assert torch.equal(outs[0][3], outs[1][3]), 'the residual passes the gradient through'
    for a, c in zip(outs[0][1:3], outs[1][1:3]):
        assert _rel(a, c) <= TOL


def conv_backend_stats():
    from gan_track_b200.torch_utils.ops import conv_backend
    return conv_backend.stats
