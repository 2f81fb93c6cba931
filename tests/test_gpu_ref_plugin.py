"""GPU (-m gpu): the C-ABI kernels against the REFERENCE'S OWN CUDA PLUGIN on identical device inputs.

oracle/build_ref_plugin.py compiles OPS/bias_act.{cpp,cu} and OPS/upfirdn2d.{cpp,cu} from /root/reference for sm_100a into
oracle/_ref/ (shipped to the GPU box like the product library).  Both sides are called through the same pybind-style entry
points (OPS/bias_act.cpp:94-97, OPS/upfirdn2d.cpp:102-105): the reference module and `custom_ops.get_plugin(...)`.
North star: fp32 within 1e-5 relative, fp16 layers within 1e-2; the bounds asserted here are tighter and the achieved error is
printed.  fp16 is only reachable on the GPU in the reference (its CPU path forces fp32), so this file is the fp16 parity proof
for the two plugin ops on every call shape of SURVEY section 3.4 / Appendix B."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = 'cuda'
_CACHE = {}


def plugins():
    if 'p' not in _CACHE:
        from oracle import build_ref_plugin
        if not build_ref_plugin.built():
            pytest.skip('oracle/_ref/*.so not built (run python oracle/build_ref_plugin.py in the build container)')
        from gan_track_b200.torch_utils import custom_ops
        _CACHE['p'] = (build_ref_plugin.load(), (custom_ops.get_plugin('bias_act_plugin'), custom_ops.get_plugin('upfirdn2d_plugin')))
    return _CACHE['p']


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


TOL = {torch.float32: 2e-6, torch.float16: 1e-3}          # asserted; north star: 1e-5 / 1e-2

# (act code, alpha, gain, clamp, bias?)   activation codes: OPS/bias_act.py:21-31 `cuda_idx`
PATH_PARAMS = {
    'lrelu_sqrt2_c256': (3, 0.2, math.sqrt(2), 256.0, True),
    'lrelu_g1_c181': (3, 0.2, 1.0, 181.02, True),
    'linear_c256': (1, 0.0, 1.0, 256.0, True),
    'linear_sqrthalf_nobias': (1, 0.0, math.sqrt(0.5), -1.0, False),
    'lrelu_fc_noclamp': (3, 0.2, math.sqrt(2), -1.0, True),
}
# call shapes of the 256^2 trace at batch 4
PATH_SHAPES = [
    ('lrelu_sqrt2_c256', (4, 512, 4, 4), torch.float32, False),
    ('lrelu_sqrt2_c256', (4, 512, 16, 16), torch.float32, False),
    ('lrelu_sqrt2_c256', (4, 512, 32, 32), torch.float16, False),
    ('lrelu_sqrt2_c256', (4, 512, 32, 32), torch.float16, True),
    ('lrelu_sqrt2_c256', (4, 256, 64, 64), torch.float16, True),
    ('lrelu_sqrt2_c256', (4, 128, 128, 128), torch.float16, True),
    ('lrelu_sqrt2_c256', (4, 64, 256, 256), torch.float16, False),
    ('lrelu_sqrt2_c256', (4, 64, 256, 256), torch.float16, True),
    ('lrelu_g1_c181', (4, 128, 128, 128), torch.float16, True),
    ('lrelu_g1_c181', (4, 512, 8, 8), torch.float32, False),
    ('linear_c256', (4, 1, 256, 256), torch.float32, False),
    ('linear_c256', (4, 1, 32, 32), torch.float16, False),
    ('linear_sqrthalf_nobias', (4, 128, 128, 128), torch.float16, True),
    ('linear_sqrthalf_nobias', (4, 512, 16, 16), torch.float32, False),
    ('lrelu_fc_noclamp', (32, 512), torch.float32, False),
    ('lrelu_sqrt2_c256', (3, 40, 7, 5), torch.float16, False),          # ragged: no vector width divides it
]


@pytest.mark.parametrize('case', PATH_SHAPES, ids=lambda c: f'{c[0]}-{"x".join(map(str, c[1]))}-{str(c[2])[6:]}-{"cl" if c[3] else "nchw"}')
def test_bias_act_vs_reference_plugin_path_shapes(case):
    (ref_ba, _), (our_ba, _) = plugins()
    pname, shape, dtype, cl = case
    act, alpha, gain, clamp, has_b = PATH_PARAMS[pname]
    g = torch.Generator(device=DEV).manual_seed(hash(case[0]) % 1000 + len(shape))
    x = (torch.randn(shape, device=DEV, generator=g) * 3).to(dtype)
    dy = torch.randn(shape, device=DEV, generator=g).to(dtype)
    ddx = torch.randn(shape, device=DEV, generator=g).to(dtype)
    if cl:
        x, dy, ddx = (v.contiguous(memory_format=torch.channels_last) for v in (x, dy, ddx))
    b = torch.randn([shape[1]], device=DEV, generator=g).to(dtype) if has_b else torch.empty([0], device=DEV, dtype=dtype)
    e = torch.empty([0], device=DEV, dtype=dtype)
    # grad 0 (forward), grad 1 (first derivative given the saved OUTPUT, OPS/bias_act.py:175-183), grad 1 again on d_dx (the
    # R1 / path-length double backward of linear / lrelu, OPS/bias_act.py:194-203)
    y_ref = ref_ba.bias_act(x, b, e, e, e, 0, 1, act, alpha, gain, clamp)
    y_our = our_ba.bias_act(x, b, e, e, e, 0, 1, act, alpha, gain, clamp)
    assert y_our.shape == y_ref.shape and y_our.stride() == y_ref.stride() and y_our.dtype == dtype
    errs = [rel(y_our, y_ref)]
    yref = y_ref if act == 3 else e
    for src in (dy, ddx):
        d_ref = ref_ba.bias_act(src, e, e, yref, e, 1, 1, act, alpha, gain, clamp)        # linear / lrelu save only y (OPS/bias_act.py:151-154)
        d_our = our_ba.bias_act(src, e, e, yref, e, 1, 1, act, alpha, gain, clamp)
        errs.append(rel(d_our, d_ref))
    print(f'\n  bias_act {pname} {shape} {dtype}: y {errs[0]:.1e} dx {errs[1]:.1e} d_dy {errs[2]:.1e}; bit-identical y: {bool(torch.equal(y_our, y_ref))}')
    assert max(errs) <= TOL[dtype], errs


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('act', range(1, 10), ids=['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish'])
def test_bias_act_vs_reference_plugin_all_activations(act, dtype):
    """All nine activations, grad 0 / 1 / 2, with the defaults of OPS/bias_act.py:21-31; `ref` letters say what each one saves."""
    (ref_ba, _), (our_ba, _) = plugins()
    table = {1: (0.0, 1.0, ''), 2: (0.0, math.sqrt(2), 'y'), 3: (0.2, math.sqrt(2), 'y'), 4: (0.0, 1.0, 'y'), 5: (0.0, 1.0, 'y'),
             6: (0.0, 1.0, 'y'), 7: (0.0, 1.0, 'y'), 8: (0.0, 1.0, 'y'), 9: (0.0, math.sqrt(2), 'x')}
    alpha, gain, saves = table[act]
    g = torch.Generator(device=DEV).manual_seed(act)
    shape = (4, 48, 20, 12)
    x = (torch.randn(shape, device=DEV, generator=g) * 2).to(dtype)
    b = torch.randn([48], device=DEV, generator=g).to(dtype)
    dy = torch.randn(shape, device=DEV, generator=g).to(dtype)
    v = torch.randn(shape, device=DEV, generator=g).to(dtype)
    e = torch.empty([0], device=DEV, dtype=dtype)
    tol = {torch.float32: 1e-5, torch.float16: 2e-3}[dtype]       # transcendental activations: the reference is built with --use_fast_math
    for clamp in (-1.0, 1.5):
        y_ref = ref_ba.bias_act(x, b, e, e, e, 0, 1, act, alpha, gain, clamp)
        y_our = our_ba.bias_act(x, b, e, e, e, 0, 1, act, alpha, gain, clamp)
        assert rel(y_our, y_ref) <= tol, ('y', act, clamp, rel(y_our, y_ref))
        keep_x = saves == 'x' or act >= 4            # x and b are saved when 'x' in ref or has_2nd_grad (OPS/bias_act.py:151-154)
        xref, bs = (x, b) if keep_x else (e, e)
        yref = y_ref if saves == 'y' else e
        d_ref = ref_ba.bias_act(dy, bs, xref, yref, e, 1, 1, act, alpha, gain, clamp)
        d_our = our_ba.bias_act(dy, bs, xref, yref, e, 1, 1, act, alpha, gain, clamp)
        assert rel(d_our, d_ref) <= tol, ('grad1', act, clamp, rel(d_our, d_ref))
        if act >= 4:            # has_2nd_grad (OPS/bias_act.py:24-31)
            s_ref = ref_ba.bias_act(v, bs, xref, yref, dy, 2, 1, act, alpha, gain, clamp)
            s_our = our_ba.bias_act(v, bs, xref, yref, dy, 2, 1, act, alpha, gain, clamp)
            assert rel(s_our, s_ref) <= 5 * tol, ('grad2', act, clamp, rel(s_our, s_ref))


F4 = np.outer([1, 3, 3, 1], [1, 3, 3, 1]).astype(np.float32) / 64
SYM6 = np.asarray([0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466, 0.787641141030194,
                   0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036,
                   -0.007800708325034148], dtype=np.float32)
SYM6 = SYM6 / SYM6.sum()
# (name, filter [fh, fw], (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain), x shape)  -- SURVEY Appendix B
UP_CASES = [
    ('g_blur_after_convT', F4, (1, 1, 1, 1, 1, 1, 1, 1, False, 4.0), (4, 64, 65, 65)),
    ('g_blur_after_convT_257', F4, (1, 1, 1, 1, 1, 1, 1, 1, False, 4.0), (2, 64, 257, 257)),
    ('g_blur_512ch', F4, (1, 1, 1, 1, 1, 1, 1, 1, False, 4.0), (4, 512, 33, 33)),
    ('d_conv1_preblur', F4, (1, 1, 1, 1, 2, 2, 2, 2, False, 1.0), (4, 64, 64, 64)),
    ('d_conv1_preblur_256', F4, (1, 1, 1, 1, 2, 2, 2, 2, False, 1.0), (2, 64, 256, 256)),
    ('g_img_up2', F4, (2, 2, 1, 1, 2, 1, 2, 1, False, 4.0), (4, 1, 128, 128)),
    ('d_skip_down2', F4, (1, 1, 2, 2, 1, 1, 1, 1, False, 1.0), (4, 64, 256, 256)),
    ('d_skip_down2_bwd', F4, (2, 2, 1, 1, 2, 1, 2, 1, True, 1.0), (4, 64, 128, 128)),          # its backward: up 2, flipped (OPS/upfirdn2d.py:256-266)
    ('g_img_up2_bwd', F4, (1, 1, 2, 2, 1, 2, 1, 2, True, 4.0), (4, 1, 256, 256)),
    ('aug_up2_x', SYM6[None, :], (2, 1, 1, 1, 6, 5, 0, 0, False, 2.0), (4, 1, 140, 150)),
    ('aug_up2_y', SYM6[:, None], (1, 2, 1, 1, 0, 0, 6, 5, False, 2.0), (4, 1, 140, 300)),
    ('aug_down2_x_flip', SYM6[None, :], (1, 1, 2, 1, -1, -1, 0, 0, True, 1.0), (4, 1, 140, 300)),
    ('aug_down2_y_flip', SYM6[:, None], (1, 1, 1, 2, 0, 0, -1, -1, True, 1.0), (4, 1, 280, 145)),
    ('generic_large_kernel', np.random.RandomState(0).rand(7, 5).astype(np.float32), (3, 2, 2, 3, 4, 1, 0, 5, False, 0.7), (2, 3, 19, 23)),
    ('crop_negative_pad', F4, (1, 1, 1, 1, -1, 2, 0, -2, False, 1.0), (2, 6, 33, 40)),
]


@pytest.mark.parametrize('cl', [False, True], ids=['nchw', 'cl'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('case', UP_CASES, ids=lambda c: c[0])
def test_upfirdn2d_vs_reference_plugin(case, dtype, cl):
    (_, ref_up), (_, our_up) = plugins()
    name, f, args, shape = case
    if cl and shape[1] == 1:
        pytest.skip('single channel: layouts coincide')
    g = torch.Generator(device=DEV).manual_seed(len(name))
    x = torch.randn(shape, device=DEV, generator=g).to(dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    ft = torch.from_numpy(np.ascontiguousarray(f)).to(DEV)
    y_ref = ref_up.upfirdn2d(x, ft, *args)
    y_our = our_up.upfirdn2d(x, ft, *args)
    assert y_our.shape == y_ref.shape and y_our.dtype == y_ref.dtype
    assert y_our.stride() == y_ref.stride(), 'output memory format must follow x.suggest_memory_format() (OPS/upfirdn2d.cpp:38)'
    e = rel(y_our, y_ref)
    print(f'\n  upfirdn2d {name} {shape} {dtype} {"cl" if cl else "nchw"}: {e:.1e}')
    assert e <= {torch.float32: 2e-6, torch.float16: 1e-3}[dtype], e
