"""FromRGB with one image channel (csrc/rgb.cu): `fromrgb1(x, w, b, act, gain, clamp)` =
`bias_act(conv2d(x, w.reshape(C, 1, 1, 1)), b, act, gain, clamp)` for x of shape [N, 1, H, W], in one pass and written
channels-last -- Conv2dLayer(img_channels=1 -> C, kernel 1) of the discriminator's top block
(S3/training/networks_stylegan2.py:586, 617-621).

Backward is one fused pass when no higher-order graph is being recorded; under `create_graph` (R1 differentiates the
discriminator's backward, S3/training/loss.py:120-133) the same formulas run as differentiable tensor ops on top of
bias_act's own gradient Function."""
import torch

from ... import _lib
from . import bias_act as bias_act_mod


def applicable(x, out_channels):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.ndim == 4 and x.shape[1] == 1 and x.dtype in (torch.float16, torch.float32)):
        return False
    vec = 8 if x.dtype == torch.float16 else 4
    cv = out_channels // vec
    return out_channels % vec == 0 and 1 <= cv <= 32 and (cv & (cv - 1)) == 0 and x.numel() > 0


_cache = {}


def _fn(act, alpha, gain, clamp):
    bias_act_mod._init()
    BA = bias_act_mod._bias_act_cuda(dim=1, act=act, alpha=alpha, gain=gain, clamp=clamp)
    if BA in _cache:
        return _cache[BA]
    spec, alpha_f, gain_f, clamp_f, trivial = BA.cfg

    class FromRGB1(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b):
            N, _, H, W = x.shape
            C = w.numel()
            xc = x.contiguous()
            wc = w.to(x.dtype).contiguous()
            bc = b.to(x.dtype).contiguous() if b is not None else None
            y = torch.empty([N, C, H, W], dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().gt_fromrgb1_fwd(_lib.ptr(xc), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(y), _lib.dtype_code(x), spec.cuda_idx, alpha_f,
                                                       gain_f, clamp_f, N * H * W, C, _lib.stream_of(x)), 'gt_fromrgb1_fwd')
            _lib.count_launch()
            ctx.save_for_backward(x, w, y)
            ctx.has_bias = b is not None
            return y

        @staticmethod
        def backward(ctx, dy):
            x, w, y = ctx.saved_tensors
            N, _, H, W = x.shape
            C = w.numel()
            need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
            dy = dy.contiguous(memory_format=torch.channels_last)
            if torch.is_grad_enabled():
                # differentiable form: bias_act's gradient Function (from the saved output), then the 1x1 convolution's gradients
                e = bias_act_mod._empty
                g1 = dy if trivial else BA.Grad.apply(dy, e, e, y if 'y' in spec.ref else e)
                wv = w.to(dy.dtype).reshape(1, C, 1, 1)
                dx = (g1 * wv).sum(dim=1, keepdim=True) if need_x else None
                dw = (g1 * x).sum(dim=[0, 2, 3]).to(w.dtype).reshape(w.shape) if need_w else None
                db = g1.sum(dim=[0, 2, 3]) if need_b else None
                return dx, dw, db
            lib = _lib.load()
            # `linear` saves no output in the reference, so its clamp does not mask the gradient (OPS/bias_act.py:151-154)
            clamp_bwd = clamp_f if 'y' in spec.ref else -1.0
            dx = torch.empty_like(x, memory_format=torch.contiguous_format) if need_x else None
            dw = torch.empty([C], dtype=torch.float32, device=x.device)
            db = torch.empty([C], dtype=torch.float32, device=x.device)
            nws = lib.gt_fromrgb1_bwd_workspace(C)
            ws = torch.empty([nws], dtype=torch.float32, device=x.device)
            xc, wc = x.contiguous(), w.to(x.dtype).contiguous()
            with torch.cuda.device(x.device):
                _lib.check(lib.gt_fromrgb1_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(xc), _lib.ptr(wc), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), nws,
                                               _lib.dtype_code(x), spec.cuda_idx, alpha_f, gain_f, clamp_bwd, N * H * W, C, _lib.stream_of(x)), 'gt_fromrgb1_bwd')
            _lib.count_launch(2)
            return dx, (dw.to(w.dtype).reshape(w.shape) if need_w else None), (db.to(x.dtype) if need_b else None)

    _cache[BA] = FromRGB1
    return FromRGB1


def fromrgb1(x, w, b=None, act='linear', alpha=None, gain=None, clamp=None):
    """x: [N,1,H,W]; w: [C] (any shape with C elements: the 1x1 kernel); b: [C] or None -> [N,C,H,W] channels-last."""
    _lib.require_cuda(x, 'fromrgb input')
    return _fn(act, alpha, gain, clamp).apply(x, w, b)


# ---- ToRGB with one image channel ---------------------------------------------------------------------------------------

torgb_fused = True      # switched off around passes that differentiate the generator's backward (see loss._g_pathlen)


class op_by_op_torgb:
    """Context: ToRGB takes the op-by-op route.  Under `create_graph` the fused op would have to re-run that route inside its
    backward anyway (measured: +3 ms per path-length pass), so the loss switches it off for that pass up front."""

    def __enter__(self):
        global torgb_fused
        self.old, torgb_fused = torgb_fused, False

    def __exit__(self, *a):
        global torgb_fused
        torgb_fused = self.old


def torgb_applicable(x, out_channels):
    if not torgb_fused:
        return False
    if out_channels != 1 or not (isinstance(x, torch.Tensor) and x.is_cuda and x.ndim == 4 and x.dtype in (torch.float16, torch.float32)):
        return False
    n, c, h, w = x.shape
    vec = 8 if x.dtype == torch.float16 else 4
    cv = c // vec
    if c % vec or not (1 <= cv <= 32) or (cv & (cv - 1)) or n == 0 or n > 65535 or h * w == 0:
        return False
    st = x.stride()
    return st[1] == 1 and st[3] == c and st[2] == w * c and st[0] == h * w * c and x.data_ptr() % 16 == 0


def _torgb_reference(x, weight, styles, bias, clamp):
    """The op-by-op form (S3/training/networks_stylegan2.py:68-77 with demodulate=False, then :357): differentiable to any order."""
    from . import conv2d_gradfix
    n = x.shape[0]
    xs = x * styles.to(x.dtype).reshape(n, -1, 1, 1)
    y = conv2d_gradfix.conv2d(xs, weight.to(x.dtype))
    return bias_act_mod.bias_act(y, bias.to(x.dtype) if bias is not None else None, clamp=clamp)


class _ToRGB1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, styles, bias, clamp):
        N, C, H, W = x.shape
        wc = weight.to(x.dtype).reshape(-1).contiguous()
        sc = styles.to(torch.float32).contiguous()
        bc = bias.to(x.dtype).contiguous() if bias is not None else None
        y = torch.empty([N, 1, H, W], dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().gt_torgb1_fwd(_lib.ptr(x), _lib.ptr(sc), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(y), _lib.dtype_code(x),
                                                 float(clamp) if clamp is not None else -1.0, N, H * W, C, _lib.stream_of(x)), 'gt_torgb1_fwd')
        _lib.count_launch()
        ctx.save_for_backward(x, weight, styles, bias)
        ctx.clamp = clamp
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, styles, bias = ctx.saved_tensors
        N, C, H, W = x.shape
        if torch.is_grad_enabled():
            # a higher-order graph is being recorded (path-length regulariser): differentiate the op-by-op form instead
            ins = [t for t in (x, weight, styles, bias) if t is not None and t.requires_grad]
            with torch.enable_grad():
                y = _torgb_reference(x, weight, styles, bias, ctx.clamp)
                grads = torch.autograd.grad([y], ins, [dy], create_graph=True, allow_unused=True)
            it = iter(grads)
            return tuple((next(it) if (t is not None and t.requires_grad) else None) for t in (x, weight, styles, bias)) + (None,)
        lib = _lib.load()
        dyc = dy.contiguous()
        wc = weight.to(x.dtype).reshape(-1).contiguous()
        sc = styles.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        ds = torch.empty([N, C], dtype=torch.float32, device=x.device)
        dw = torch.empty([C], dtype=torch.float32, device=x.device)
        db = torch.empty([1], dtype=torch.float32, device=x.device)
        nws = lib.gt_torgb1_bwd_workspace(N, C)
        ws = torch.empty([nws], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.gt_torgb1_bwd(_lib.ptr(dyc), _lib.ptr(x), _lib.ptr(sc), _lib.ptr(wc), _lib.ptr(dx), _lib.ptr(ds), _lib.ptr(dw), _lib.ptr(db),
                                         _lib.ptr(ws), nws, _lib.dtype_code(x), N, H * W, C, _lib.stream_of(x)), 'gt_torgb1_bwd')
        _lib.count_launch(2)
        return (dx if ctx.needs_input_grad[0] else None, dw.to(weight.dtype).reshape(weight.shape) if ctx.needs_input_grad[1] else None,
                ds.to(styles.dtype) if ctx.needs_input_grad[2] else None,
                db.to(bias.dtype).reshape(bias.shape) if (bias is not None and ctx.needs_input_grad[3]) else None, None)


def torgb1(x, weight, styles, bias=None, clamp=None):
    """x: [N,C,H,W] channels-last; weight: [1,C,1,1]; styles: [N,C]; bias: [1] or None  ->  [N,1,H,W]."""
    _lib.require_cuda(x, 'torgb input')
    return _ToRGB1.apply(x, weight, styles, bias, clamp)
