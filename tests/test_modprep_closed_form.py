"""CPU, float64: the closed-form first and second derivative of the style side of the operand preparation (oracle/modprep_ref.py; the
formulas csrc/modprep.cu evaluates on the GPU) against autograd through the reference's op chain
(S3/training/networks_stylegan2.py:52-63)."""
import pytest
import torch

from oracle import modprep_ref as R


@pytest.mark.parametrize('prenorm', [True, False])
@pytest.mark.parametrize('with_a', [True, False])
def test_style_side_closed_form_matches_autograd(prenorm, with_a):
    torch.manual_seed(3)
    dt = torch.float64
    N, I, O = 6, 20, 9
    s = (torch.randn(N, I, dtype=dt) * 1.5 + 0.5).requires_grad_(True)
    wsq = (torch.rand(O, I, dtype=dt) + 0.05).requires_grad_(True)
    a = torch.randn(N, I, dtype=dt).requires_grad_(True) if with_a else None
    b = torch.randn(N, O, dtype=dt).requires_grad_(True)
    u = torch.randn(N, I, dtype=dt)
    sn, d = R.chain(s, wsq, prenorm)
    outs, cots = ([sn, d], [a, b]) if with_a else ([d], [b])
    gs_ref, gw_ref = torch.autograd.grad(outs, [s, wsq], cots, create_graph=True)
    gs, gw, _ = R.first_order(a, b, s, wsq, prenorm)
    assert (gs - gs_ref).abs().max() <= 1e-12 and (gw - gw_ref).abs().max() <= 1e-12
    wrt = ([a] if with_a else []) + [b, s, wsq]
    ref = list(torch.autograd.grad(gs_ref, wrt, u))
    det = lambda t: t.detach() if t is not None else None
    gga, ggb, g2s, g2w = R.second_order(u, det(a), det(b), det(s), det(wsq), prenorm)
    if with_a:
        assert (gga - ref.pop(0)).abs().max() <= 1e-12
    for got, want in zip((ggb, g2s, g2w), ref):
        assert (got - want).abs().max() <= 1e-11 * max(1.0, float(want.abs().max()))
