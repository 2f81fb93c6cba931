"""Whole-model parity at widths the tensor-core kernels cover, against the REAL reference (tests/golden/model_wide.npz,
written by oracle/gen_golden_wide.py: resolution 64, 64 / 128 channels, all four loss phases with ADA, style mixing and
both lazy regularisers on one device-independent random stream).

  * CPU: the product's host code over the oracle ops reproduces the reference (pins draw order and host logic at this size);
  * GPU fp32: every block in fp32 on the CUDA path, tolerance 1e-5 of the north star on the network outputs and STATED
    per-phase gradient tolerances (printed with the achieved error);
  * GPU fp16: the default configuration (blocks at resolution >= 8 in fp16 -> tcgen05 implicit-GEMM convolutions) within the
    north star's 1e-2 on the outputs, with the convolution routes asserted to be the repository's own kernels.

Gradient tensors are compared on the fixture's strided sample (error relative to the tensor's largest reference entry) and
through the L2 norm of the whole tensor.

WHY THE GRADIENT TOLERANCES ARE NOT 1e-5.  The network OUTPUTS are continuous in the arithmetic and are held to 1e-5 (fp32) /
1e-2 (fp16).  First-order parameter gradients are not: every leaky-ReLU gate is a step function of its pre-activation, and among
the ~2e6 pre-activations of one pass a few lie within fp32 rounding distance of zero.  Two correct fp32 implementations that
round differently (here: the CPU reference vs the same host code over the oracle ops, measured in this container) disagree on
ONE such gate in Dmain: `b32.conv1.bias` differs by 2.1e-3 of its largest entry in exactly one channel while the other 127
channels agree to 7e-7, and everything upstream of that pixel moves by ~1e-4 (Gmain: median tensor error 1.6e-4 because the
flipped gate sits in D, through which all of G's gradient flows).  Phases without such an event agree to 5e-6 (Dreg) and 5e-5
(Greg).  The reference compared with itself at another CPU thread count stays at 8e-6 only because its convolutions are then
bit-identical.  Hence three numbers per phase: the median over tensors (typical error), the worst tensor norm, and the worst
single entry (gate flips), each with its own stated bound and the achieved value printed."""
import numpy as np
import pytest
import torch

from oracle import gen_golden_wide as gw


def _run(device, G_kw=None, D_kw=None, oracle=False, loss_extra=None):
    from gan_track_b200.training import augment, loss as loss_mod, networks_stylegan2 as nets
    G, D = gw.build(nets, G_kw, D_kw)
    G, D = G.to(device), D.to(device)
    z, c, real = (t.to(device) for t in gw.make_inputs())
    if oracle:
        from oracle.backend import oracle_ops
        with oracle_ops():
            return gw.run_phases(loss_mod, augment.AugmentPipe, G, D, z, c, real, device=device, loss_extra=loss_extra)
    return gw.run_phases(loss_mod, augment.AugmentPipe, G, D, z, c, real, device=device, loss_extra=loss_extra)


def _compare(res, golden, label):
    """-> {'out': worst output error, phase: (worst sampled-entry error, worst norm error, name)}; asserts structure only."""
    Z = golden('model_wide.npz')
    worst = {}
    for k in ['G_train_random', 'D_real', 'G_eval_const']:
        ref = Z['wide/' + k]
        got = res[k].detach().float().cpu().numpy()
        assert got.shape == ref.shape, k
        worst[k] = float(np.abs(got - ref).max() / np.abs(ref).max())
    for phase, which, _ in gw.PHASES:
        e_max, n_max, who, checked, per_tensor = 0.0, 0.0, None, 0, []
        num2 = den2 = 0.0
        phase_absmax = max(float(Z[k]) for k in Z.keys(f'wide/{phase}/') if k.endswith('/absmax'))
        for key in Z.keys(f'wide/{phase}/'):
            if not key.endswith('/sample'):
                continue
            name = key[len(f'wide/{phase}/'):-len('/sample')]
            absmax = float(Z[f'wide/{phase}/{name}/absmax'])
            if absmax == 0.0:
                continue            # structurally zero second-order gradients: None on the CUDA-plugin semantics (OPS/bias_act.py:196-204)
            assert f'{phase}/{name}' in res, f'{label}: no gradient for {phase}/{name}'
            g = res[f'{phase}/{name}'].detach().float().cpu().numpy().reshape(-1)
            idx = gw.sample_index(g.size)
            e = float(np.abs(g[idx] - Z[key]).max() / absmax)
            nref = float(Z[f'wide/{phase}/{name}/norm'])
            n = abs(float(np.sqrt((g.astype(np.float64) ** 2).sum())) - nref) / nref
            num2 += float(((g[idx].astype(np.float64) - Z[key]) ** 2).sum())
            den2 += float((Z[key].astype(np.float64) ** 2).sum())
            # worst entry / norm over the tensors that matter: more than one element (a scalar such as a noise strength is one
            # cancellation-prone sum) and not orders of magnitude below the phase's largest gradient (second-order bias gradients of
            # ~1e-6 are below fp16 resolution of the tensors they flow through)
            if g.size > 1 and absmax >= 1e-3 * phase_absmax:
                if e > e_max:
                    e_max, who = e, name
                n_max = max(n_max, n)
            per_tensor.append(e)
            checked += 1
        assert checked >= 40, (phase, checked)
        worst[phase] = (e_max, n_max, who, float(np.median(per_tensor)), float(np.sqrt(num2 / den2)))
        pm = float(res[f'{phase}/pl_mean'])
        assert abs(pm - float(Z[f'wide/{phase}/pl_mean'])) <= 2e-2 * abs(float(Z[f'wide/{phase}/pl_mean'])), (phase, pm)
    print(f'\n[{label}] ' + '  '.join(f'{k}={v:.2e}' if not isinstance(v, tuple) else f'{k}: global {v[4]:.2e} median {v[3]:.2e} entry {v[0]:.2e} norm {v[1]:.2e} ({v[2]})'
                                     for k, v in worst.items()))
    return worst


def test_host_code_over_oracle_matches_reference_wide(golden):
    torch.set_num_threads(8)
    w = _compare(_run('cpu', oracle=True), golden, 'cpu host+oracle')
    for k in ['G_train_random', 'D_real', 'G_eval_const']:
        assert w[k] <= 1e-5, (k, w[k])
    _check_grads(w, FP32_GRAD_TOL)


def _check_grads(w, tol):
    for phase, _, _ in gw.PHASES:
        entry, norm, who, median, glob = w[phase]
        assert glob <= tol['global'], (phase, w[phase])
        assert median <= tol['median'], (phase, w[phase])
        if 'norm' in tol:
            assert norm <= tol['norm'], (phase, w[phase])
        if 'entry' in tol:
            assert entry <= tol['entry'], (phase, w[phase])


# Stated tolerances (see the module docstring): outputs at the north star's bounds; gradients as global L2 error over all sampled
# entries of the phase, median over tensors of the worst entry, and (fp32 only) worst norm / entry over the significant tensors.
# fp16: thousands of leaky-ReLU gates flip between an fp16 and an fp32 evaluation, so single tensors move by several percent
# (measured on B200: Gmain median 5e-2, Dmain 6e-3) -- the same holds for the library's fp16 kernels, which
# test_cuda_fp16_library_route_for_comparison prints next to ours.
FP32_OUT_TOL = 1e-5
FP32_GRAD_TOL = dict(**{'global': 5e-3}, median=5e-4, norm=1e-2, entry=1e-2)
FP16_OUT_TOL = 1e-2
FP16_GRAD_TOL = dict(**{'global': 2e-1}, median=1e-1)


@pytest.mark.gpu
def test_cuda_fp32_matches_reference_wide(golden):
    from gan_track_b200.torch_utils.ops import conv_backend
    w = _compare(_run('cuda', dict(gw.G_KW, num_fp16_res=0), dict(gw.D_KW, num_fp16_res=0)), golden, 'cuda fp32')
    for k in ['G_train_random', 'D_real', 'G_eval_const']:
        assert w[k] <= FP32_OUT_TOL, (k, w[k])
    _check_grads(w, FP32_GRAD_TOL)
    print('conv routes', conv_backend.stats)


@pytest.mark.gpu
@pytest.mark.parametrize('merged', [False, True], ids=['reference_schedule', 'merged_schedule'])
def test_cuda_fp16_tensor_core_route_matches_reference_wide(golden, merged):
    """Default configuration on a CUDA device: blocks at resolution 8..64 in fp16, every 64 / 128-channel convolution of them
    (forward, data gradient, weight gradient, and the double-backwards of R1 / path length) on the tcgen05 kernels."""
    from gan_track_b200.torch_utils.ops import conv_backend
    before = dict(conv_backend.stats)
    extra = dict(merge_d_passes=True, merge_mapping_passes=True) if merged else None
    w = _compare(_run('cuda', loss_extra=extra), golden, 'cuda fp16' + (' merged' if merged else ''))
    used = {k: conv_backend.stats[k] - before[k] for k in before}
    print('conv routes of this run', used)
    assert used['igemm'] >= 100 and used['igemm_wgrad'] >= 40, used
    for k in ['G_train_random', 'D_real', 'G_eval_const']:
        assert w[k] <= FP16_OUT_TOL, (k, w[k])
    _check_grads(w, FP16_GRAD_TOL)


@pytest.mark.gpu
def test_cuda_fp16_library_route_for_comparison(golden):
    """The same fp16 model with every convolution on the LIBRARY (what the reference itself calls, OPS/conv2d_gradfix.py:37-45):
    printed next to the tensor-core route above to show that the fp16 gradient deviations from the fp32 golden are a property of
    fp16 arithmetic, not of this package's kernels; only the output bound is asserted."""
    from gan_track_b200.torch_utils.ops import conv_backend
    old = conv_backend.allow_igemm
    conv_backend.allow_igemm = False
    try:
        w = _compare(_run('cuda'), golden, 'cuda fp16, library convolutions')
    finally:
        conv_backend.allow_igemm = old
    for k in ['G_train_random', 'D_real', 'G_eval_const']:
        assert w[k] <= FP16_OUT_TOL, (k, w[k])
