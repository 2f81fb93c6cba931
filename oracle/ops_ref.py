"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the arithmetic of Gan-track's StyleGAN2-ADA op surface.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import this.

Every function states the behaviour of one reference op and cites the reference lines it follows
(S3 = /root/reference/src/models/stylegan3, OPS = S3/torch_utils/ops).  The arithmetic the reference
delegates to PyTorch (convolution, grid_sample, matmul: OPS/conv2d_gradfix.py:40,45,
OPS/grid_sample_gradfix.py:31) is delegated to the same PyTorch CPU kernels here, because that IS the
reference's algorithm on its CPU (`impl='ref'`) path.

Parity pin: `oracle/gen_golden.py` imports the real reference in the build container and stores its outputs
on seeded inputs under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against
those vectors.  The reference itself ships no tests or golden vectors for this path.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------
# bias_act  (OPS/bias_act.py:21-31 table, :91-120 ref path, OPS/bias_act.cu:23-146 cuda semantics)
# ---------------------------------------------------------------------------------------------------

_SELU_SCALE = 1.0507009873554804934193349852946
_SELU_ALPHA = 1.6732632423543772848170429916717

# name -> (cuda_idx, def_alpha, def_gain, ref, has_2nd_grad)      OPS/bias_act.py:21-31
ACTIVATIONS = {
    'linear':   (1, 0.0, 1.0,          '',  False),
    'relu':     (2, 0.0, math.sqrt(2), 'y', False),
    'lrelu':    (3, 0.2, math.sqrt(2), 'y', False),
    'tanh':     (4, 0.0, 1.0,          'y', True),
    'sigmoid':  (5, 0.0, 1.0,          'y', True),
    'elu':      (6, 0.0, 1.0,          'y', True),
    'selu':     (7, 0.0, 1.0,          'y', True),
    'softplus': (8, 0.0, 1.0,          'y', True),
    'swish':    (9, 0.0, math.sqrt(2), 'x', True),
}


def _act_fwd(u, act, alpha):
    if act == 'linear':
        return u
    if act == 'relu':
        return torch.relu(u)
    if act == 'lrelu':
        return torch.where(u > 0, u, u * alpha)
    if act == 'tanh':
        return torch.tanh(u)
    if act == 'sigmoid':
        return torch.sigmoid(u)
    if act == 'elu':
        return F.elu(u)
    if act == 'selu':
        return F.selu(u)
    if act == 'softplus':
        return F.softplus(u)
    if act == 'swish':
        return torch.sigmoid(u) * u
    raise KeyError(act)


def _resolve(act, alpha, gain, clamp):
    _, def_alpha, def_gain, _, _ = ACTIVATIONS[act]
    alpha = float(def_alpha if alpha is None else alpha)
    gain = float(def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    return alpha, gain, clamp


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='ref'):
    """y = clamp(act(x + b) * gain).  Follows OPS/bias_act.py:91-120 (differentiable through torch autograd)."""
    alpha, gain, clamp = _resolve(act, alpha, gain, clamp)
    if b is not None:
        assert b.ndim == 1 and b.shape[0] == x.shape[dim]
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = _act_fwd(x, act, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def bias_act_kernel(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp):
    """Restatement of the CUDA kernel semantics for grad in {0,1,2} (OPS/bias_act.cu:39-142), computed in the
    kernel's internal precision (fp32 for fp16/fp32 I/O, fp64 for fp64) and cast back.  `x` is the differentiated
    quantity: the input (grad=0), dy (grad=1) or d_dx (grad=2); `b` is added to x (grad=0) or xref (grad>0)."""
    io = x.dtype
    ct = torch.float64 if io == torch.float64 else torch.float32
    shp = [-1 if i == dim else 1 for i in range(x.ndim)]
    X = x.to(ct)
    B = b.to(ct).reshape(shp) if b is not None and b.numel() else 0
    XR = xref.to(ct) if xref is not None and xref.numel() else torch.zeros((), dtype=ct)
    YR = yref.to(ct) if yref is not None and yref.numel() else torch.zeros((), dtype=ct)
    DY = dy.to(ct) if dy is not None and dy.numel() else torch.ones((), dtype=ct)
    yy = YR / gain if gain != 0 else torch.zeros_like(YR)
    if grad == 0:
        X = X + B
    else:
        XR = XR + B
    one = 1.0
    if act == 'linear':
        y = X if grad < 2 else torch.zeros_like(X)
    elif act == 'relu':
        y = {0: lambda: torch.where(X > 0, X, 0 * X), 1: lambda: torch.where(yy > 0, X, 0 * X), 2: lambda: 0 * X}[grad]()
    elif act == 'lrelu':
        y = {0: lambda: torch.where(X > 0, X, X * alpha), 1: lambda: torch.where(yy > 0, X, X * alpha), 2: lambda: 0 * X}[grad]()
    elif act == 'tanh':
        y = {0: lambda: torch.tanh(X), 1: lambda: X * (one - yy * yy), 2: lambda: X * (one - yy * yy) * (-2 * yy)}[grad]()
    elif act == 'sigmoid':
        y = {0: lambda: torch.sigmoid(X), 1: lambda: X * yy * (one - yy), 2: lambda: X * yy * (one - yy) * (one - 2 * yy)}[grad]()
    elif act == 'elu':
        y = {0: lambda: torch.where(X >= 0, X, torch.expm1(X)), 1: lambda: torch.where(yy >= 0, X, X * (yy + one)),
             2: lambda: torch.where(yy >= 0, 0 * X, X * (yy + one))}[grad]()
    elif act == 'selu':
        sa = _SELU_SCALE * _SELU_ALPHA
        y = {0: lambda: torch.where(X >= 0, _SELU_SCALE * X, sa * torch.expm1(X)),
             1: lambda: torch.where(yy >= 0, X * _SELU_SCALE, X * (yy + sa)),
             2: lambda: torch.where(yy >= 0, 0 * X, X * (yy + sa))}[grad]()
    elif act == 'softplus':
        def g2():
            c = torch.exp(-yy)
            return X * c * (one - c)
        y = {0: lambda: F.softplus(X), 1: lambda: X * (one - torch.exp(-yy)), 2: g2}[grad]()
    elif act == 'swish':
        if grad == 0:
            y = X * torch.sigmoid(X)
        else:
            c = torch.exp(XR)
            d = c + one
            if grad == 1:
                y = torch.where(XR > 40, X, X * c * (XR + d) / (d * d))
            else:
                y = torch.where(XR > 40, 0 * X, X * c * (XR * (2 - d) + 2 * d) / (d * d * d))
            YR = XR * torch.sigmoid(XR) * gain
    else:
        raise KeyError(act)
    y = y * (gain * DY)
    if clamp >= 0:
        if grad == 0:
            y = y.clamp(-clamp, clamp)
        else:
            y = torch.where((YR > -clamp) & (YR < clamp), y, torch.zeros_like(y))
    return y.to(io)


# ---------------------------------------------------------------------------------------------------
# upfirdn2d  (OPS/upfirdn2d.py:70-114 setup_filter, :167-211 ref path, :251-266 backward rule, :277-387 wrappers)
# ---------------------------------------------------------------------------------------------------

def _scaling(s):
    if isinstance(s, int):
        s = [s, s]
    sx, sy = s
    assert sx >= 1 and sy >= 1
    return int(sx), int(sy)


def _padding(p):
    if isinstance(p, int):
        p = [p, p]
    p = list(p)
    if len(p) == 2:
        p = [p[0], p[0], p[1], p[1]]
    return tuple(int(v) for v in p)


def _filter_size(f):
    if f is None:
        return 1, 1
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """OPS/upfirdn2d.py:70-114: fp32, optional outer product for <8 taps, normalise to unit DC, flip, gain."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='ref'):
    """Zero-insert upsample, pad/crop, FIR, decimate.  OPS/upfirdn2d.py:167-211 (differentiable via autograd)."""
    assert x.ndim == 4
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
    n, c, h, w = x.shape
    upx, upy = _scaling(up)
    downx, downy = _scaling(down)
    px0, px1, py0, py1 = _padding(padding)
    assert w * upx + px0 + px1 >= f.shape[-1] and h * upy + py0 + py1 >= f.shape[0]
    x = x.reshape(n, c, h, 1, w, 1)
    x = F.pad(x, [0, upx - 1, 0, 0, 0, upy - 1])
    x = x.reshape(n, c, h * upy, w * upx)
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    f = f * (gain ** (f.ndim / 2))
    f = f.to(x.dtype)
    if not flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f[None, None].repeat([c, 1] + [1] * f.ndim)
    if f.ndim == 4:
        x = F.conv2d(x, f, groups=c)
    else:
        x = F.conv2d(x, f.unsqueeze(2), groups=c)
        x = F.conv2d(x, f.unsqueeze(3), groups=c)
    return x[:, :, ::downy, ::downx]


def upfirdn2d_direct(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Second, independent statement of the same op as explicit loops over taps (the form the
    CUDA kernels implement: OPS/upfirdn2d.cu:29-92).  numpy float64; used to cross-check `upfirdn2d` on small cases."""
    x = np.asarray(x, dtype=np.float64)
    f = np.ones([1, 1]) if f is None else np.asarray(f, dtype=np.float64)
    if f.ndim == 1:
        f = np.outer(f, f)
    n, c, h, w = x.shape
    upx, upy = _scaling(up)
    downx, downy = _scaling(down)
    px0, px1, py0, py1 = _padding(padding)
    fh, fw = f.shape
    oh = (h * upy + py0 + py1 - fh + downy) // downy
    ow = (w * upx + px0 + px1 - fw + downx) // downx
    g = f if flip_filter else f[::-1, ::-1]
    y = np.zeros([n, c, oh, ow])
    for ky in range(fh):
        for kx in range(fw):
            for oy in range(oh):
                py = oy * downy + ky - py0
                if py < 0 or py % upy or py // upy >= h:
                    continue
                for ox in range(ow):
                    qx = ox * downx + kx - px0
                    if qx < 0 or qx % upx or qx // upx >= w:
                        continue
                    y[:, :, oy, ox] += g[ky, kx] * x[:, :, py // upy, qx // upx]
    return y * gain


def upfirdn2d_backward_args(x_shape, y_shape, f, up, down, padding, flip_filter, gain):
    """Arguments of the op that IS the gradient w.r.t. x (OPS/upfirdn2d.py:251-266)."""
    upx, upy = _scaling(up)
    downx, downy = _scaling(down)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    _, _, ih, iw = x_shape
    _, _, oh, ow = y_shape
    p = [fw - px0 - 1, iw * upx - ow * downx + px0 - upx + 1, fh - py0 - 1, ih * upy - oh * downy + py0 - upy + 1]
    return dict(up=[downx, downy], down=[upx, upy], padding=p, flip_filter=(not flip_filter), gain=gain)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='ref'):
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='ref'):
    upx, upy = _scaling(up)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='ref'):
    downx, downy = _scaling(down)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


# ---------------------------------------------------------------------------------------------------
# conv2d / conv_transpose2d / conv2d_resample  (OPS/conv2d_gradfix.py:37-45, OPS/conv2d_resample.py:29-141)
# ---------------------------------------------------------------------------------------------------

def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(input, weight, bias, stride, padding, dilation, groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    return F.conv_transpose2d(input, weight, bias, stride, padding, output_padding, groups, dilation)


def _conv(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    kh, kw = w.shape[2:]
    if not flip_weight and (kw > 1 or kh > 1):     # OPS/conv2d_resample.py:35-37
        w = w.flip([2, 3])
    if transpose:
        return F.conv_transpose2d(x, w, stride=stride, padding=padding, groups=groups)
    return F.conv2d(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """Same case split as OPS/conv2d_resample.py:82-141."""
    oc, icg, kh, kw = w.shape
    fw, fh = _filter_size(f)
    px0, px1, py0, py1 = _padding(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    if kw == 1 and kh == 1 and down > 1 and up == 1:
        x = upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:
        x = _conv(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:
        x = upfirdn2d(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, oc // groups, icg, kh, kw).transpose(1, 2).reshape(groups * icg, oc // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv(x, w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return _conv(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    x = upfirdn2d(x, (f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = _conv(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x


# ---------------------------------------------------------------------------------------------------
# fma, grid_sample  (OPS/fma.py:15-58, OPS/grid_sample_gradfix.py:23-31)
# ---------------------------------------------------------------------------------------------------

def fma(a, b, c):
    return torch.addcmul(c, a, b)


def grid_sample(input, grid):
    return F.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros', align_corners=False)


def grid_sample_direct(img, grid):
    """Independent statement of bilinear / zeros / align_corners=False sampling in numpy float64."""
    img = np.asarray(img, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float64)
    n, c, h, w = img.shape
    _, oh, ow, _ = grid.shape
    ix = ((grid[..., 0] + 1) * w - 1) / 2
    iy = ((grid[..., 1] + 1) * h - 1) / 2
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    out = np.zeros([n, c, oh, ow])
    for dy in (0, 1):
        for dx in (0, 1):
            xx, yy = x0 + dx, y0 + dy
            wgt = (1 - np.abs(ix - xx)) * (1 - np.abs(iy - yy))
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            xc, yc = np.clip(xx, 0, w - 1), np.clip(yy, 0, h - 1)
            for b in range(n):
                out[b] += img[b][:, yc[b], xc[b]] * (wgt[b] * ok[b])[None]
    return out


# ---------------------------------------------------------------------------------------------------
# modulated_conv2d  (S3/training/networks_stylegan2.py:32-89)
# ---------------------------------------------------------------------------------------------------

def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    n = x.shape[0]
    oc, ic, kh, kw = weight.shape
    if x.dtype == torch.float16 and demodulate:                                  # :52-54
        weight = weight * (1 / np.sqrt(ic * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    w = None
    dcoefs = None
    if demodulate or fused_modconv:                                              # :59-61
        w = weight.unsqueeze(0) * styles.reshape(n, 1, -1, 1, 1)
    if demodulate:                                                               # :63
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(n, -1, 1, 1, 1)
    if not fused_modconv:                                                        # :68-77
        x = x * styles.to(x.dtype).reshape(n, -1, 1, 1)
        x = conv2d_resample(x, weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = fma(x, dcoefs.to(x.dtype).reshape(n, -1, 1, 1), noise.to(x.dtype))
        elif demodulate:
            x = x * dcoefs.to(x.dtype).reshape(n, -1, 1, 1)
        elif noise is not None:
            x = x + noise.to(x.dtype)
        return x
    x = x.reshape(1, -1, *x.shape[2:])                                            # :79-89
    w = w.reshape(-1, ic, kh, kw)
    x = conv2d_resample(x, w.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, groups=n, flip_weight=flip_weight)
    x = x.reshape(n, -1, *x.shape[2:])
    if noise is not None:
        x = x + noise
    return x


def dcoefs_closed_form(weight, styles):
    """d[n,o] = rsqrt(sum_i s[n,i]^2 * sum_k W[o,i,k]^2 + 1e-8)  (equals networks_stylegan2.py:60-63)."""
    wsq = weight.double().square().sum(dim=[2, 3])          # [O,I]
    return (styles.double().square() @ wsq.t() + 1e-8).rsqrt()
