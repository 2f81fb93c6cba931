"""Fused element-wise halves of the training-mode modulated convolution (csrc/modulated.cu) on channels-last tensors.

    mod_scale(x, s)                         = x * s.to(x.dtype)[:, :, None, None]               networks_stylegan2.py:69
    demod_act(x, d, noise, b, act, ...)     = bias_act(fma(x, d.to(x.dtype)[:, :, None, None], noise), b, act, gain, clamp)
                                                                                              networks_stylegan2.py:71-72, 325-327
Each is ONE pass over the activation instead of one per torch op; each backward is ONE pass that also produces every
reduction (style / demodulation-coefficient / noise / bias gradients) in a fixed order.  Both first-order backward
functions are themselves differentiable (the path-length regulariser differentiates the generator's backward,
S3/training/loss.py:85-100): their backward is again ONE fused pass (gt_mod_scale_bwd2 / gt_demod_act_bwd2) when no
higher-order graph is being recorded, and the same formulas written with tensor ops otherwise.

`applicable(x)` says whether a tensor can take the fused route (CUDA, fp16/fp32, dense channels-last, channel count the
kernels support); callers keep the reference's op-by-op sequence otherwise.
"""
import numpy as np
import torch

from ... import _lib

_ACT = {'linear': 1, 'lrelu': 3}


def applicable(x):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.ndim == 4 and x.dtype in (torch.float16, torch.float32)):
        return False
    n, c, h, w = x.shape
    vec = 8 if x.dtype == torch.float16 else 4
    if c % vec or n == 0 or h * w == 0:
        return False
    cv = c // vec
    if not ((cv < 32 and 256 % cv == 0) or (cv >= 32 and cv % 32 == 0 and cv <= 128)):
        return False
    s = x.stride()
    return s[1] == 1 and s[3] == c and s[2] == w * c and s[0] == h * w * c and x.data_ptr() % 16 == 0


def _cl(t):
    return t if applicable(t) else t.contiguous(memory_format=torch.channels_last)


def _ws(x):
    n, c, h, w = x.shape
    lib = _lib.load()
    nws = lib.gt_mod_workspace(n, h * w, c, _lib.dtype_code(x))
    return torch.empty([nws], dtype=torch.float32, device=x.device), nws


def _bc(v, like):
    """[N,C] fp32 -> [N,C,1,1] in like.dtype"""
    return v.to(like.dtype).reshape(v.shape[0], v.shape[1], 1, 1)


# ------------------------------------------------------------------------------------------------------- mod_scale

class _ModScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        n, c, h, w = x.shape
        s = s.to(torch.float32).contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().gt_mod_scale_fwd(_lib.ptr(x), _lib.ptr(s), _lib.ptr(y), _lib.dtype_code(x), n, h * w, c, _lib.stream_of(x)), 'gt_mod_scale_fwd')
        _lib.count_launch()
        ctx.save_for_backward(x, s)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, s = ctx.saved_tensors
        gx, gs = _ModScaleGrad.apply(gy, x, s)
        return (gx if ctx.needs_input_grad[0] else None), (gs if ctx.needs_input_grad[1] else None)


class _ModScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, x, s):
        gy = _cl(gy)
        n, c, h, w = x.shape
        gx = torch.empty_like(x)
        gs = torch.empty([n, c], dtype=torch.float32, device=x.device)
        ws, nws = _ws(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().gt_mod_scale_bwd(_lib.ptr(gy), _lib.ptr(x), _lib.ptr(s), _lib.ptr(gx), _lib.ptr(gs), _lib.ptr(ws), nws,
                                                    _lib.dtype_code(x), n, h * w, c, _lib.stream_of(x)), 'gt_mod_scale_bwd')
        _lib.count_launch(2)
        ctx.save_for_backward(gy, x, s)
        return gx, gs

    @staticmethod
    def backward(ctx, ggx, ggs):
        # gx = gy * s,  gs = sum_hw gy * x   =>   second-order terms below
        gy, x, s = ctx.saved_tensors
        d_gy = d_x = d_s = None
        if not torch.is_grad_enabled() and applicable(gy) and applicable(x) and (ggx is not None or ggs is not None):
            # one fused pass (csrc/modulated.cu: mod_scale_bwd2_kernel); the tensor-op formulas below are the differentiable form
            n, c, h, w = x.shape
            ggx_c = _cl(ggx) if ggx is not None else None
            ggs_c = ggs.to(torch.float32).contiguous() if ggs is not None else None
            if ctx.needs_input_grad[0]:
                d_gy = torch.empty_like(x)
            if ctx.needs_input_grad[1] and ggs is not None:
                d_x = torch.empty_like(x)
            if ctx.needs_input_grad[2] and ggx is not None:
                d_s = torch.empty([n, c], dtype=torch.float32, device=x.device)
            ws, nws = _ws(x)
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().gt_mod_scale_bwd2(_lib.ptr(ggx_c), _lib.ptr(ggs_c), _lib.ptr(gy), _lib.ptr(x), _lib.ptr(s), _lib.ptr(d_gy), _lib.ptr(d_x),
                                                         _lib.ptr(d_s), _lib.ptr(ws), nws, _lib.dtype_code(x), n, h * w, c, _lib.stream_of(x)), 'gt_mod_scale_bwd2')
            _lib.count_launch(2 if d_s is not None else 1)
            return d_gy, d_x, d_s
        if ctx.needs_input_grad[0]:
            d_gy = 0
            if ggx is not None:
                d_gy = d_gy + ggx * _bc(s, x)
            if ggs is not None:
                d_gy = d_gy + _bc(ggs, x) * x
        if ctx.needs_input_grad[1] and ggs is not None:
            d_x = _bc(ggs, x) * gy
        if ctx.needs_input_grad[2] and ggx is not None:
            d_s = (ggx.to(torch.float32) * gy.to(torch.float32)).sum(dim=[2, 3])
        return d_gy, d_x, d_s


def mod_scale(x, s):
    """x: [N,C,H,W] channels-last; s: [N,C] (any float dtype; applied in x.dtype as the reference does)."""
    return _ModScale.apply(x, s)


# ------------------------------------------------------------------------------------------------------- demod_act

def _slope(y, act, alpha, gain, clamp):
    """gain * act'(.) * [|y| < clamp], decided from the saved output like the reference plugin (OPS/bias_act.cu:132-142)."""
    m = torch.full_like(y, gain)
    if act == 'lrelu':
        m = torch.where(y > 0, m, m * alpha)
    if clamp >= 0:
        m = m * (y.abs() < clamp).to(y.dtype)
    return m


class _DemodAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, d, noise, b, act, alpha, gain, clamp):
        n, c, h, w = x.shape
        d32 = d.to(torch.float32).contiguous() if d is not None else None
        nz = noise.to(x.dtype).expand(n, 1, h, w).contiguous() if noise is not None else None
        bb = b.to(x.dtype).contiguous() if b is not None else None
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().gt_demod_act_fwd(_lib.ptr(x), _lib.ptr(d32), _lib.ptr(nz), _lib.ptr(bb), _lib.ptr(y), _lib.dtype_code(x), _ACT[act],
                                                    alpha, gain, clamp, n, h * w, c, _lib.stream_of(x)), 'gt_demod_act_fwd')
        _lib.count_launch()
        ctx.save_for_backward(x if d is not None else None, d32, y)
        ctx.cfg = (act, alpha, gain, clamp, noise is not None, tuple(noise.shape) if noise is not None else None, b is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, d, y = ctx.saved_tensors
        act, alpha, gain, clamp, has_nz, nz_shape, has_b = ctx.cfg
        gx, gd, gnz, s0 = _DemodActGrad.apply(gy, y, x, d, act, alpha, gain, clamp, has_nz and ctx.needs_input_grad[2])
        g_noise = None
        if has_nz and ctx.needs_input_grad[2]:
            g_noise = gnz
            if tuple(gnz.shape) != nz_shape:                     # noise was broadcast (e.g. [H,W] constant noise)
                g_noise = gnz.sum_to_size(nz_shape)
        g_b = s0.sum(dim=0).to(y.dtype) if (has_b and ctx.needs_input_grad[3]) else None
        return (gx if ctx.needs_input_grad[0] else None), (gd if (d is not None and ctx.needs_input_grad[1]) else None), g_noise, g_b, None, None, None, None


class _DemodActGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, y, x, d, act, alpha, gain, clamp, want_noise):
        gy = _cl(gy)
        n, c, h, w = y.shape
        gx = torch.empty_like(y)
        gd = torch.empty([n, c], dtype=torch.float32, device=y.device) if d is not None else None
        s0 = torch.empty([n, c], dtype=torch.float32, device=y.device)
        gnz = torch.empty([n, 1, h, w], dtype=y.dtype, device=y.device) if want_noise else None
        ws, nws = _ws(y)
        with torch.cuda.device(y.device):
            _lib.check(_lib.load().gt_demod_act_bwd(_lib.ptr(gy), _lib.ptr(y), _lib.ptr(x), _lib.ptr(d), _lib.ptr(gx), _lib.ptr(gnz), _lib.ptr(gd), _lib.ptr(s0),
                                                    _lib.ptr(ws), nws, _lib.dtype_code(y), _ACT[act], alpha, gain, clamp, n, h * w, c, _lib.stream_of(y)),
                       'gt_demod_act_bwd')
        _lib.count_launch(3 if d is not None else 2)
        ctx.save_for_backward(gy, y, x, d)
        ctx.cfg = (act, alpha, gain, clamp)
        return gx, gd, gnz, s0

    @staticmethod
    def backward(ctx, ggx, ggd, ggnz, ggs0):
        # With m = gain * act'(y) * [|y| < clamp] (piecewise constant) and g1 = gy * m:
        #   gx = g1 * d,  gd = sum_hw g1 * x,  gnz = sum_c g1,  s0 = sum_hw g1
        gy, y, x, d = ctx.saved_tensors
        act, alpha, gain, clamp = ctx.cfg
        if not torch.is_grad_enabled() and applicable(gy) and applicable(y) and any(t is not None for t in (ggx, ggd, ggnz, ggs0)):
            # one fused pass (csrc/modulated.cu: demod_act_bwd2_kernel); the tensor-op formulas below are the differentiable form
            n, c, h, w = y.shape
            ggx_c = _cl(ggx) if ggx is not None else None
            ggd_c = ggd.to(torch.float32).contiguous() if (ggd is not None and d is not None) else None
            ggs0_c = ggs0.to(torch.float32).contiguous() if ggs0 is not None else None
            ggnz_c = ggnz.to(y.dtype).expand(n, 1, h, w).contiguous() if ggnz is not None else None
            d_gy = torch.empty_like(y) if ctx.needs_input_grad[0] else None
            d_x = torch.empty_like(y) if (d is not None and ctx.needs_input_grad[2] and ggd_c is not None) else None
            d_d = torch.empty([n, c], dtype=torch.float32, device=y.device) if (d is not None and ctx.needs_input_grad[3] and ggx is not None) else None
            if d_gy is None and d_x is None and d_d is None:
                return None, None, None, None, None, None, None, None, None
            ws, nws = _ws(y)
            with torch.cuda.device(y.device):
                _lib.check(_lib.load().gt_demod_act_bwd2(_lib.ptr(ggx_c), _lib.ptr(ggd_c), _lib.ptr(ggnz_c), _lib.ptr(ggs0_c), _lib.ptr(gy), _lib.ptr(y),
                                                         _lib.ptr(x if d is not None else None), _lib.ptr(d), _lib.ptr(d_gy), _lib.ptr(d_x), _lib.ptr(d_d),
                                                         _lib.ptr(ws), nws, _lib.dtype_code(y), _ACT[act], alpha, gain, clamp, n, h * w, c,
                                                         _lib.stream_of(y)), 'gt_demod_act_bwd2')
            _lib.count_launch(2 if d_d is not None else 1)
            return d_gy, None, d_x, d_d, None, None, None, None, None
        m = _slope(y, act, alpha, gain, clamp)
        d_gy = d_x = d_d = None
        if ctx.needs_input_grad[0]:
            t = 0
            if ggx is not None:
                t = t + (ggx * _bc(d, y) if d is not None else ggx)
            if ggd is not None and d is not None:
                t = t + _bc(ggd, y) * x
            if ggnz is not None:
                t = t + ggnz.to(y.dtype)
            if ggs0 is not None:
                t = t + _bc(ggs0, y)
            d_gy = m * t if not isinstance(t, int) else None
        g1 = None
        if d is not None and ctx.needs_input_grad[2] and ggd is not None:
            g1 = gy * m
            d_x = _bc(ggd, y) * g1
        if d is not None and ctx.needs_input_grad[3] and ggx is not None:
            g1 = gy * m if g1 is None else g1
            d_d = (ggx.to(torch.float32) * g1.to(torch.float32)).sum(dim=[2, 3])
        return d_gy, None, d_x, d_d, None, None, None, None, None


def demod_act(x, d=None, noise=None, b=None, act='linear', alpha=0.2, gain=1.0, clamp=None):
    """clamp(act(x * d[n,c] + noise + b[c]) * gain) on a channels-last tensor.  d: [N,C] or None; noise: broadcastable to
    [N,1,H,W] or None; b: [C] or None; act in {'linear', 'lrelu'}; clamp None or < 0 disables clamping."""
    assert act in _ACT
    return _DemodAct.apply(x, d, noise, b, act, float(alpha), float(gain), float(clamp) if clamp is not None else -1.0)


# ---------------------------------------------------------------------------------------------------------------------
# parameter / style side: pre-normalisation + demodulation coefficients (csrc/modprep.cu)
# ---------------------------------------------------------------------------------------------------------------------

def prep_tensor_ops(weight, styles, prenorm):
    """The same three results as `prep` as plain tensor ops (S3/training/networks_stylegan2.py:52-63 without the [N,O,I,kh,kw]
    product): differentiable to any order by autograd.  Used as the second-order form of `_ModPrep.backward`."""
    if prenorm:
        fan_in = weight.shape[1] * weight.shape[2] * weight.shape[3]
        weight = weight * (1 / np.sqrt(fan_in) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    d = (styles.square() @ weight.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()
    return (weight.to(torch.float16), styles, d) if prenorm else (None, None, d)


fused_prep = True          # False: callers keep the tensor-op chain (tests compare the two routes)
# which form evaluated the second-order pass of the style side (tests assert the path-length pattern stays on the closed form)
prep_stats = {'style_bwd2_closed_form': 0, 'style_bwd2_autograd': 0}


def prep_applicable(weight, styles):
    """The fused route is closed under differentiation to second order in the styles (what the path-length pass needs: closed-form
    kernels, `_PrepStyleGrad.backward`) and falls back to autograd over the tensor-op form for anything beyond (a cotangent on the
    weight-side gradient, third order), so it is used wherever the shapes fit."""
    return (fused_prep and weight.is_cuda and weight.dtype == torch.float32 and styles.dtype == torch.float32 and styles.ndim == 2
            and weight.ndim == 4 and 1 <= styles.shape[0] <= 64 and weight.shape[1] % 4 == 0 and styles.shape[1] == weight.shape[1])


def _weight_chain(weight, prenorm):
    """Weight side of `prep_tensor_ops`: (scaled fp16 weight | None, wsq [O, I])."""
    if prenorm:
        fan_in = weight.shape[1] * weight.shape[2] * weight.shape[3]
        weight = weight * (1 / np.sqrt(fan_in) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
    wsq = weight.square().sum(dim=[2, 3])
    return (weight.to(torch.float16) if prenorm else None), wsq


def _style_chain(styles, wsq, prenorm):
    """Style side of `prep_tensor_ops`: (normalised styles | the styles, dcoefs [N, O])."""
    sn = styles / styles.norm(float('inf'), dim=1, keepdim=True) if prenorm else styles
    return sn, (sn.square() @ wsq.t() + 1e-8).rsqrt()


def _leaf(t):
    return t.detach().requires_grad_(True)


class _PrepWeight(torch.autograd.Function):
    """weight [O,I,kh,kw] fp32 -> (fp16 pre-normalised weight | empty, wsq [O,I]); csrc/modprep.cu, first-order kernels.  The path-length
    pass never differentiates this backward (the weight is not on a path to the latents); if some other caller does, the backward is
    evaluated through autograd on the tensor-op form."""

    @staticmethod
    def forward(ctx, weight, prenorm):
        lib = _lib.load()
        O, I, kh, kw = weight.shape
        W = weight.contiguous()
        dev = W.device
        wsq = torch.empty([O, I], dtype=torch.float32, device=dev)
        w16 = scale = amax = None
        if prenorm:
            w16 = torch.empty([O, I, kh, kw], dtype=torch.float16, device=dev)
            scale = torch.empty([O], dtype=torch.float32, device=dev)
            amax = torch.empty([O], dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.gt_modprep_weight_fwd(_lib.ptr(W), _lib.ptr(w16), _lib.ptr(wsq), _lib.ptr(scale), _lib.ptr(amax), O, I, kh * kw, int(prenorm),
                                                 _lib.stream_of(W)), 'gt_modprep_weight_fwd')
        _lib.count_launch()
        ctx.save_for_backward(W, scale, amax, weight)
        ctx.prenorm = bool(prenorm)
        ctx.set_materialize_grads(False)
        if w16 is None:
            w16 = torch.empty([0], dtype=torch.float16, device=dev)
            ctx.mark_non_differentiable(w16)
        return w16, wsq

    @staticmethod
    def backward(ctx, g_w, g_wsq):
        W, scale, amax, weight_in = ctx.saved_tensors
        if not ctx.needs_input_grad[0] or (g_w is None and g_wsq is None):
            return None, None
        if torch.is_grad_enabled():          # only inside backward(create_graph=True) with the weight on the differentiated path
            with torch.enable_grad():
                wi = weight_in if weight_in.requires_grad else _leaf(weight_in)
                w16, wsq = _weight_chain(wi, ctx.prenorm)
                pairs = [(o, g) for o, g in ((w16, g_w if ctx.prenorm else None), (wsq, g_wsq)) if o is not None and g is not None]
                gW, = torch.autograd.grad([o for o, _ in pairs], [wi], [g.to(o.dtype) for o, g in pairs], create_graph=True)
            return gW, None
        O, I, kh, kw = W.shape
        if g_wsq is None:
            g_wsq = torch.zeros([O, I], dtype=torch.float32, device=W.device)
        g_wsq = g_wsq.contiguous().to(torch.float32)
        g_w = g_w.contiguous() if (g_w is not None and ctx.prenorm) else None
        gW = torch.empty_like(W)
        with torch.cuda.device(W.device):
            _lib.check(_lib.load().gt_modprep_weight_bwd(_lib.ptr(W), _lib.ptr(g_w), _lib.dtype_code(g_w) if g_w is not None else 0, _lib.ptr(g_wsq),
                                                         _lib.ptr(scale), _lib.ptr(amax), _lib.ptr(gW), O, I, kh * kw, int(ctx.prenorm), _lib.stream_of(W)),
                       'gt_modprep_weight_bwd')
        _lib.count_launch()
        return gW, None


class _PrepStyle(torch.autograd.Function):
    """(styles [N,I], wsq [O,I]) -> (normalised styles | empty, dcoefs [N,O]); backward = `_PrepStyleGrad`, itself differentiable."""

    @staticmethod
    def forward(ctx, styles, wsq, prenorm):
        from . import fc
        lib = _lib.load()
        s, wsq_c = styles.contiguous(), wsq.contiguous()
        N, I = s.shape
        dev = s.device
        f32 = dict(dtype=torch.float32, device=dev)
        sn2 = torch.empty([N, I], **f32)
        smax = sarg = None
        sn = s
        if prenorm:
            sn = torch.empty([N, I], **f32)
            smax = torch.empty([N], **f32)
            sarg = torch.empty([N], dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_of(s)
            _lib.check(lib.gt_modprep_style_fwd(_lib.ptr(s), _lib.ptr(sn) if prenorm else None, _lib.ptr(sn2), _lib.ptr(smax), _lib.ptr(sarg), N, I,
                                                int(prenorm), st), 'gt_modprep_style_fwd')
            q = fc._fwd(sn2, wsq_c, None, 1.0, 0.0)                                # [N, O]
            d = torch.empty_like(q)
            _lib.check(lib.gt_modprep_rsqrt(_lib.ptr(q), _lib.ptr(d), q.numel(), 1e-8, st), 'gt_modprep_rsqrt')
        _lib.count_launch(2)
        ctx.save_for_backward(styles, wsq, sn, sn2, smax, sarg, d)
        ctx.prenorm = bool(prenorm)
        ctx.set_materialize_grads(False)
        if prenorm:
            return sn, d
        empty = torch.empty([0], **f32)
        ctx.mark_non_differentiable(empty)
        return empty, d

    @staticmethod
    def backward(ctx, g_sn, g_d):
        styles, wsq, sn, sn2, smax, sarg, d = ctx.saved_tensors
        if not ctx.prenorm:
            g_sn = None
        if g_sn is None and g_d is None:
            return None, None, None
        if g_d is None:
            g_d = torch.zeros_like(d)
        gs, g_wsq = _PrepStyleGrad.apply(g_sn, g_d, styles, wsq, sn, sn2, smax, sarg, d, ctx.prenorm)
        return (gs if ctx.needs_input_grad[0] else None), (g_wsq if ctx.needs_input_grad[1] else None), None


class _PrepStyleGrad(torch.autograd.Function):
    """(g_sn | None, g_d, styles, wsq) -> (gs, g_wsq): the first-order vector-Jacobian product of `_PrepStyle` on the fused kernels.  Its own
    backward (a cotangent on gs: the path-length pass) is closed form, csrc/modprep.cu `style_bwd2_*`."""

    @staticmethod
    def forward(ctx, g_sn, g_d, styles, wsq, sn, sn2, smax, sarg, d, prenorm):
        from . import fc
        lib = _lib.load()
        N, I = sn.shape
        g_d = g_d.contiguous().to(torch.float32)
        g_sn = g_sn.contiguous().to(torch.float32) if g_sn is not None else None
        wsq_c = wsq.contiguous()
        gq = torch.empty_like(d)
        gs = torch.empty_like(sn)
        with torch.cuda.device(sn.device):
            st = _lib.stream_of(sn)
            _lib.check(lib.gt_modprep_gq(_lib.ptr(d), _lib.ptr(g_d), _lib.ptr(gq), d.numel(), st), 'gt_modprep_gq')
            gp = fc._dgrad(gq, wsq_c, 1.0)                                         # [N, I]
            _lib.check(lib.gt_modprep_style_bwd(_lib.ptr(g_sn), _lib.ptr(sn), _lib.ptr(gp), _lib.ptr(smax), _lib.ptr(sarg), _lib.ptr(gs), N, I, int(prenorm), st),
                       'gt_modprep_style_bwd')
            g_wsq = fc._wgrad(gq, sn2, 1.0, 0.0, False)[0]                         # [O, I]
        _lib.count_launch(2)
        ctx.save_for_backward(g_sn, g_d, styles, wsq, sn, sn2, smax, sarg, d, gp)
        ctx.prenorm = bool(prenorm)
        ctx.set_materialize_grads(False)         # an unused g_wsq must arrive as None in backward, not as a tensor of zeros
        return gs, g_wsq

    @staticmethod
    def backward(ctx, u, V):
        from . import fc
        g_sn, g_d, styles, wsq, sn, sn2, smax, sarg, d, gp = ctx.saved_tensors
        prenorm = ctx.prenorm
        need = ctx.needs_input_grad
        none6 = (None,) * 6
        if u is None and V is None:
            return (None, None, None, None) + none6
        if V is not None or torch.is_grad_enabled():
            # beyond what the closed form covers (cotangent on the weight-side gradient, third order): autograd over the tensor-op form
            prep_stats['style_bwd2_autograd'] += 1
            with torch.enable_grad():
                a = _leaf(g_sn) if g_sn is not None else None
                b, s_, w_ = _leaf(g_d), _leaf(styles), _leaf(wsq)
                o_sn, o_d = _style_chain(s_, w_, prenorm)
                outs, cots = ([o_sn, o_d], [a, b]) if a is not None else ([o_d], [b])
                gs_, gw_ = torch.autograd.grad(outs, [s_, w_], cots, create_graph=True)
                pairs = [(o, c) for o, c in ((gs_, u), (gw_, V)) if c is not None]
                wrt = [t for t in (a, b, s_, w_) if t is not None]
                got = list(torch.autograd.grad([o for o, _ in pairs], wrt, [c for _, c in pairs], allow_unused=True))
            gga = got.pop(0) if a is not None else None
            ggb, g2s, g2w = got
            return (gga if need[0] else None, ggb if need[1] else None, g2s if need[2] else None, g2w if need[3] else None) + none6
        prep_stats['style_bwd2_closed_form'] += 1
        lib = _lib.load()
        N, I = sn.shape
        O = d.shape[1]
        dev = sn.device
        f32 = dict(dtype=torch.float32, device=dev)
        u = u.contiguous().to(torch.float32)
        wsq_c = wsq.contiguous()
        v, vs, r = torch.empty([N, I], **f32), torch.empty([N, I], **f32), torch.empty([N], **f32)
        ggd = torch.empty([N, O], **f32)
        g2 = torch.empty([2 * N, O], **f32)            # rows [gq ; hq]
        x2 = torch.empty([2 * N, I], **f32)            # rows [2 v sn / m ; sn^2]
        gga = torch.empty([N, I], **f32) if (g_sn is not None and need[0]) else None
        g2s = torch.empty([N, I], **f32)
        with torch.cuda.device(dev):
            st = _lib.stream_of(sn)
            _lib.check(lib.gt_modprep_style_bwd2_a(_lib.ptr(u), _lib.ptr(sn), _lib.ptr(sarg), _lib.ptr(v), _lib.ptr(vs), _lib.ptr(r), N, I, int(prenorm), st),
                       'gt_modprep_style_bwd2_a')
            z = fc._fwd(vs, wsq_c, None, 1.0, 0.0)                                 # [N, O]
            _lib.check(lib.gt_modprep_style_bwd2_b(_lib.ptr(d), _lib.ptr(g_d), _lib.ptr(z), _lib.ptr(smax), _lib.ptr(ggd), _lib.ptr(g2[:N]), _lib.ptr(g2[N:]),
                                                   N, O, int(prenorm), st), 'gt_modprep_style_bwd2_b')
            hp = fc._dgrad(g2[N:], wsq_c, 1.0)                                     # [N, I]
            _lib.check(lib.gt_modprep_style_bwd2_c(_lib.ptr(g_sn), _lib.ptr(sn), _lib.ptr(gp), _lib.ptr(hp), _lib.ptr(v), _lib.ptr(r), _lib.ptr(smax),
                                                   _lib.ptr(sarg), _lib.ptr(gga), _lib.ptr(g2s), _lib.ptr(x2[:N]), _lib.ptr(x2[N:]), N, I, int(prenorm), st),
                       'gt_modprep_style_bwd2_c')
            g2w = None
            if need[3]:
                if 2 * N <= 64:
                    g2w = fc._wgrad(g2, x2, 1.0, 0.0, False)[0]                    # gq^T @ x1 + hq^T @ sn^2 in one product over 2N rows
                else:
                    g2w = fc._wgrad(g2[:N], x2[:N], 1.0, 0.0, False)[0] + fc._wgrad(g2[N:], x2[N:], 1.0, 0.0, False)[0]
        _lib.count_launch(3)
        return (gga, ggd if need[1] else None, g2s if need[2] else None, g2w) + none6


def prep(weight, styles, prenorm):
    """(weight_scaled_fp16 | None, styles_normalised | None, dcoefs): the pre-normalised operands (fp16 layers) and the
    demodulation coefficients of modulated_conv2d (S3/training/networks_stylegan2.py:52-63), differentiable w.r.t. weight
    and styles; twice in the styles on the fused kernels."""
    w16, wsq = _PrepWeight.apply(weight, bool(prenorm))
    sn, d = _PrepStyle.apply(styles, wsq, bool(prenorm))
    return (w16, sn, d) if prenorm else (None, None, d)
