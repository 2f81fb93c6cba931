"""ORACLE (test infrastructure): the closed-form second derivative of the style side of modulated_conv2d's operand preparation,
written with plain tensor ops so that it can be checked against autograd in float64 on the CPU.  csrc/modprep.cu
(`modprep_style_bwd2_{a,b,c}` around three fully-connected products) evaluates exactly these formulas on the GPU.

Chain (S3/training/networks_stylegan2.py:52-63, style side):   sn = s / max|s|   (fp16 layers only; sn = s otherwise)
                                                               d  = rsqrt(sn^2 @ wsq^T + 1e-8)
Nothing under gan_track_b200/ imports this module.
"""
import torch


def chain(s, wsq, prenorm):
    sn = s / s.norm(float('inf'), dim=1, keepdim=True) if prenorm else s
    return sn, (sn.square() @ wsq.t() + 1e-8).rsqrt()


def first_order(a, b, s, wsq, prenorm):
    """Vector-Jacobian product of `chain` for cotangents a (of sn; may be None) and b (of d): (gs, g_wsq) and the intermediates."""
    if prenorm:
        m, k = s.abs().max(dim=1, keepdim=True)
        sig = torch.sign(s.gather(1, k))
        sn = s / m
    else:
        m, k, sig, sn = torch.ones_like(s[:, :1]), None, None, s
    p = sn.square()
    d = (p @ wsq.t() + 1e-8).rsqrt()
    gq = -0.5 * d ** 3 * b
    gp = gq @ wsq
    tt = (a if a is not None else 0) + 2 * sn * gp
    gs = tt / m
    if prenorm:
        gs = gs - torch.zeros_like(s).scatter_(1, k, sig * (tt * sn).sum(1, keepdim=True) / m)
    return gs, gq.t() @ p, dict(m=m, k=k, sig=sig, sn=sn, p=p, d=d, gq=gq, gp=gp, tt=tt)


def second_order(u, a, b, s, wsq, prenorm):
    """Gradient of <u, gs(a, b, s, wsq)> w.r.t. (a, b, s, wsq) in closed form."""
    _, _, c = first_order(a, b, s, wsq, prenorm)
    m, k, sig, sn, p, d, gq, gp, tt = (c[x] for x in ['m', 'k', 'sig', 'sn', 'p', 'd', 'gq', 'gp', 'tt'])
    r = sig * u.gather(1, k) if prenorm else torch.zeros_like(m)
    v = u - r * sn
    z = (v * sn) @ wsq.t()
    gga = v / m
    ggb = -d ** 3 * z / m
    hq = 1.5 * d ** 5 * b * z / m
    hp = hq @ wsq
    D = (2 * gp * v - r * tt) / m + 2 * sn * hp
    if prenorm:
        L = (tt * v).sum(1, keepdim=True) / m
        dm = -L / m - (D * sn).sum(1, keepdim=True) / m
        g2s = D / m + torch.zeros_like(s).scatter_(1, k, sig * dm)
    else:
        g2s = D
    g2w = gq.t() @ (2 * v * sn / m) + hq.t() @ p
    return gga, ggb, g2s, g2w
