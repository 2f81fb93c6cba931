"""Fused geometric resampling of the ADA pipe: reflect-pad -> 2x upsample (separable low-pass) -> bilinear
`grid_sample` on an affine grid with device-resident margins (csrc/augment_warp.cu): two passes through a workspace
(separable upsampling, then 4-tap resampling), or one 49-tap gather kernel when `two_pass` is off.

Replaces the op sequence of S3/training/augment_mi.py:303-318 (F.pad 'reflect', upfirdn2d.upsample2d,
F.affine_grid, grid_sample_gradfix.grid_sample) without the device->host read of the margins (:299).
The op is linear in the image: backward is the adjoint kernel and the double backward (R1) is the forward kernel again.
"""
import ctypes

import torch

from ... import _lib


two_pass = True      # False: the single 49-tap gather kernel (no workspace)


def _call(fn_name, src, theta, margins, taps, dst, B, C, H, W, OH, OW):
    lib = _lib.load()
    arr = (ctypes.c_float * len(taps))(*taps)
    with torch.cuda.device(src.device):
        ws, nws = None, 0
        if two_pass:
            # the 2x-upsampled image (or its gradient) at the largest possible margins; only the part the actual margins
            # need is touched
            nws = lib.gt_aug_warp_workspace(B, C, H, W)
            ws = torch.empty([nws], dtype=torch.float32, device=src.device)
        _lib.check(getattr(lib, fn_name)(_lib.ptr(src), _lib.ptr(theta), _lib.ptr(margins), arr, len(taps), _lib.ptr(dst), B, C, H, W, OH, OW,
                                         _lib.ptr(ws), nws, _lib.stream_of(src)), fn_name)
    _lib.count_launch((2 if fn_name == 'gt_aug_warp_fwd' else 3) if two_pass else 1)


def _check(theta, margins, B):
    assert theta.dtype == torch.float32 and tuple(theta.shape) == (B, 2, 3) and theta.is_contiguous()
    assert margins.dtype == torch.int32 and margins.numel() == 4 and margins.is_contiguous()


class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, theta, margins, taps, out_hw):
        _lib.require_cuda(x, 'aug_warp input')
        assert x.dtype == torch.float32 and x.ndim == 4
        x = x.contiguous()
        B, C, H, W = x.shape
        _check(theta, margins, B)
        OH, OW = out_hw
        y = torch.empty([B, C, OH, OW], dtype=torch.float32, device=x.device)
        _call('gt_aug_warp_fwd', x, theta, margins, taps, y, B, C, H, W, OH, OW)
        ctx.save_for_backward(theta, margins)
        ctx.taps, ctx.in_hw, ctx.out_hw = taps, (H, W), out_hw
        return y

    @staticmethod
    def backward(ctx, gy):
        theta, margins = ctx.saved_tensors
        gx = _WarpAdjoint.apply(gy, theta, margins, ctx.taps, ctx.in_hw) if ctx.needs_input_grad[0] else None
        return gx, None, None, None, None


class _WarpAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, theta, margins, taps, in_hw):
        gy = gy.contiguous().float()
        B, C, OH, OW = gy.shape
        _check(theta, margins, B)
        H, W = in_hw
        gx = torch.empty([B, C, H, W], dtype=torch.float32, device=gy.device)
        _call('gt_aug_warp_bwd', gy, theta, margins, taps, gx, B, C, H, W, OH, OW)
        ctx.save_for_backward(theta, margins)
        ctx.taps, ctx.out_hw = taps, (OH, OW)
        return gx

    @staticmethod
    def backward(ctx, ggx):
        theta, margins = ctx.saved_tensors
        ggy = _Warp.apply(ggx, theta, margins, ctx.taps, ctx.out_hw) if ctx.needs_input_grad[0] else None
        return ggy, None, None, None, None


def warp(x, theta, margins, taps, out_hw):
    """y[b,c] = grid_sample(upsample2d(reflect_pad(x[b,c], margins), taps, up=2), affine_grid(theta[b], out_hw)).
    x: [B,C,H,W] fp32; theta: [B,2,3] fp32 (normalised coordinates of the padded+upsampled image, as F.affine_grid takes
    them); margins: int32[4] device tensor (mx0, my0, mx1, my1), each in [0, W-1] / [0, H-1]; taps: tuple of floats
    (the normalised 1-D low-pass, even length <= 12)."""
    return _Warp.apply(x, theta, margins, tuple(float(t) for t in taps), (int(out_hw[0]), int(out_hw[1])))


def params(draws, p, multipliers, ranges, B, H, W, hz_pad):
    """All the geometric parameter algebra of the ADA pipe in one launch (csrc/augment_params.cu).  `draws`: 16 entries, (value
    draw, gate draw) per transform in the pipe's order xflip, rotate90, xint, scale, rotate (pre), aniso, rotate (post), xfrac; None
    for a disabled transform.  Returns (theta [B,2,3] fp32, margins int32 [4])."""
    lib = _lib.load()
    dev = p.device
    assert len(draws) == 16
    ptrs = (ctypes.c_void_p * 16)(*[(_lib.ptr(t.contiguous()) if t is not None else None) for t in draws])
    keep = [t for t in draws if t is not None]        # (contiguous() of a fresh rand / randn is the tensor itself)
    mult = (ctypes.c_float * 7)(*[float(m) for m in multipliers])
    rng = (ctypes.c_float * 5)(*[float(r) for r in ranges])
    theta = torch.empty([B, 2, 3], dtype=torch.float32, device=dev)
    margins = torch.empty([4], dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.gt_aug_params(ptrs, _lib.ptr(p), mult, rng, B, H, W, hz_pad, _lib.ptr(theta), _lib.ptr(margins), _lib.stream_of(theta)), 'gt_aug_params')
    _lib.count_launch()
    del keep
    return theta, margins
