"""Adaptive discriminator augmentation pipe (ADA) for single- and three-channel slices.

Interface of Gan-track's pipe (S3/training/augment_mi.py:125-195 constructor, :197-453 forward):
`AugmentPipe(run_dir, batch_size, xflip=0, rotate90=0, xint=0, ..., cutout_size=0.5)`, buffer `p` (overall strength),
`forward(images, allow_aug_debug_print=False, debug_percentile=None)`.  Random numbers are drawn in the reference's
order so a fixed seed gives the reference's augmentation parameters.

Structure: (1) per-sample parameters -> one 3x3 inverse warp `G_inv` (pixel_out -> pixel_in) and one 4x4 colour
matrix, composed here from closed-form factors; (2) execution: reflect-pad by the data-dependent margin, 2x
upsample with the 12-tap sym6 low-pass (upfirdn2d), bilinear resampling on the affine grid (grid_sample_gradfix),
2x downsample + crop (upfirdn2d); colour is a per-sample affine map; optional band filter, noise, cutout.
The debug image dump of the reference (:446-492, matplotlib) is not part of the compute path and is omitted.
"""
import numpy as np
import torch

from ..torch_utils import misc
from ..torch_utils.ops import aug_warp, conv2d_gradfix, grid_sample_gradfix, upfirdn2d

# Low-pass decomposition filters (orthogonal wavelets); the pipe uses sym6 for resampling and sym2 for the band filter.
wavelets = {
    'haar': [0.7071067811865476, 0.7071067811865476],
    'db1':  [0.7071067811865476, 0.7071067811865476],
    'db2':  [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    'db3':  [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313, 0.8068915093133388, 0.3326705529509569],
    'db4':  [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114, -0.02798376941698385, 0.6308807679295904, 0.7148465705525415, 0.23037781330885523],
    'sym2': [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    'sym3': [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313, 0.8068915093133388, 0.3326705529509569],
    'sym4': [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161, 0.29785779560527736, -0.09921954357684722, -0.012603967262037833, 0.0322231006040427],
    'sym5': [0.027333068345077982, 0.029519490925774643, -0.039134249302383094, 0.1993975339773936, 0.7234076904024206, 0.6339789634582119, 0.01660210576452232, -0.17532808990845047, -0.021101834024758855, 0.019538882735286728],
    'sym6': [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466, 0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148],
}


# ---- batched homogeneous matrices ---------------------------------------------------------------------------------

def _mat(rows, device):
    """rows: nested list mixing python scalars and tensors of a common shape S -> tensor S + [R, C]."""
    flat = [v for row in rows for v in row]
    ref = next((v for v in flat if isinstance(v, torch.Tensor)), None)
    if ref is None:
        return misc.constant(np.asarray(rows, dtype=np.float32), device=device)
    cols = [v if isinstance(v, torch.Tensor) else misc.constant(v, shape=ref.shape, device=ref.device) for v in flat]
    return torch.stack(cols, dim=-1).reshape(ref.shape + (len(rows), -1))


def translate2d(tx, ty, device=None):
    return _mat([[1, 0, tx], [0, 1, ty], [0, 0, 1]], device)


def scale2d(sx, sy, device=None):
    return _mat([[sx, 0, 0], [0, sy, 0], [0, 0, 1]], device)


def rotate2d(theta, device=None):
    return _mat([[torch.cos(theta), torch.sin(-theta), 0], [torch.sin(theta), torch.cos(theta), 0], [0, 0, 1]], device)


def translate3d(tx, ty, tz, device=None):
    return _mat([[1, 0, 0, tx], [0, 1, 0, ty], [0, 0, 1, tz], [0, 0, 0, 1]], device)


def scale3d(sx, sy, sz, device=None):
    return _mat([[sx, 0, 0, 0], [0, sy, 0, 0], [0, 0, sz, 0], [0, 0, 0, 1]], device)


def rotate3d(v, theta, device=None):
    vx, vy, vz = v[..., 0], v[..., 1], v[..., 2]
    s, c = torch.sin(theta), torch.cos(theta)
    cc = 1 - c
    return _mat([[vx * vx * cc + c, vx * vy * cc - vz * s, vx * vz * cc + vy * s, 0],
                 [vy * vx * cc + vz * s, vy * vy * cc + c, vy * vz * cc - vx * s, 0],
                 [vz * vx * cc - vy * s, vz * vy * cc + vx * s, vz * vz * cc + c, 0],
                 [0, 0, 0, 1]], device)


def translate2d_inv(tx, ty, **kw):
    return translate2d(-tx, -ty, **kw)


def scale2d_inv(sx, sy, **kw):
    return scale2d(1 / sx, 1 / sy, **kw)


def rotate2d_inv(theta, **kw):
    return rotate2d(-theta, **kw)


class AugmentPipe(torch.nn.Module):
    def __init__(self, run_dir=None, batch_size=None,
                 xflip=0, rotate90=0, xint=0, xint_max=0.125,
                 scale=0, rotate=0, aniso=0, xfrac=0, scale_std=0.2, rotate_max=1, aniso_std=0.2, xfrac_std=0.125,
                 brightness=0, contrast=0, lumaflip=0, hue=0, saturation=0, brightness_std=0.2, contrast_std=0.5, hue_max=1, saturation_std=1,
                 imgfilter=0, imgfilter_bands=[1, 1, 1, 1], imgfilter_std=1,
                 noise=0, cutout=0, noise_std=0.1, cutout_size=0.5):
        super().__init__()
        self.register_buffer('p', torch.ones([]))       # overall multiplier for augmentation probability
        self.run_dir, self.batch_size = run_dir, batch_size
        # probability multipliers and ranges, grouped as in the ADA paper
        self.xflip, self.rotate90, self.xint, self.xint_max = float(xflip), float(rotate90), float(xint), float(xint_max)
        self.scale, self.rotate, self.aniso, self.xfrac = float(scale), float(rotate), float(aniso), float(xfrac)
        self.scale_std, self.rotate_max, self.aniso_std, self.xfrac_std = float(scale_std), float(rotate_max), float(aniso_std), float(xfrac_std)
        self.brightness, self.contrast, self.lumaflip, self.hue, self.saturation = float(brightness), float(contrast), float(lumaflip), float(hue), float(saturation)
        self.brightness_std, self.contrast_std, self.hue_max, self.saturation_std = float(brightness_std), float(contrast_std), float(hue_max), float(saturation_std)
        self.imgfilter, self.imgfilter_bands, self.imgfilter_std = float(imgfilter), list(imgfilter_bands), float(imgfilter_std)
        self.noise, self.cutout, self.noise_std, self.cutout_size = float(noise), float(cutout), float(noise_std), float(cutout_size)

        self.register_buffer('Hz_geom', upfirdn2d.setup_filter(wavelets['sym6']))
        self._hz_geom_taps = tuple(float(v) for v in self.Hz_geom.tolist())     # host copy for the fused warp kernel
        self.fused_warp = True      # False: the reference's op-by-op sequence with the host read of the margins
        self.fused_params = True    # the ~110 tiny ops between the random draws and the warp as one launch (csrc/augment_params.cu)

        # Band-pass bank for the image-space filter: H(z) = sym2 low-pass, dyadic cascade of 4 bands.
        lo = np.asarray(wavelets['sym2'])
        hi = lo * ((-1) ** np.arange(lo.size))
        lo2 = np.convolve(lo, lo[::-1]) / 2
        hi2 = np.convolve(hi, hi[::-1]) / 2
        bank = np.eye(4, 1)
        for i in range(1, bank.shape[0]):
            bank = np.dstack([bank, np.zeros_like(bank)]).reshape(bank.shape[0], -1)[:, :-1]       # zero-stuff (z -> z^2)
            bank = np.stack([np.convolve(row, lo2) for row in bank])
            mid = bank.shape[1]
            bank[i, (mid - hi2.size) // 2: (mid + hi2.size) // 2] += hi2
        self.register_buffer('Hz_fbank', torch.as_tensor(bank, dtype=torch.float32))

    # -- random parameter helpers: every call consumes the generator exactly like the reference ---------------------

    def _gate(self, shape, prob, value, neutral, device):
        """value where rand(shape) < prob else neutral."""
        return torch.where(torch.rand(shape, device=device) < prob, value, neutral)

    def forward(self, images, allow_aug_debug_print=False, debug_percentile=None):
        assert isinstance(images, torch.Tensor) and images.ndim == 4
        B, C, H, W = images.shape
        dev = images.device
        dbg = None if debug_percentile is None else torch.as_tensor(debug_percentile, dtype=torch.float32, device=dev)

        # ---------------- geometric parameters -> G_inv ----------------
        I_3 = torch.eye(3, device=dev)
        G_inv = I_3
        geom_on = max(self.xflip, self.rotate90, self.xint, self.scale, self.rotate, self.aniso, self.xfrac) > 0
        if geom_on and dbg is None and self.fused_params and self.fused_warp and images.is_cuda and images.dtype == torch.float32 and B <= 1024:
            images = self._warp_fused_params(images)
            geom_on = False
        if not geom_on:
            pass
        elif self.xflip > 0:
            i = torch.floor(torch.rand([B], device=dev) * 2)
            i = self._gate([B], self.xflip * self.p, i, torch.zeros_like(i), dev)
            if dbg is not None:
                i = torch.full_like(i, torch.floor(dbg * 2))
            G_inv = G_inv @ scale2d_inv(1 - 2 * i, 1)
        if geom_on and self.rotate90 > 0:
            i = torch.floor(torch.rand([B], device=dev) * 4)
            i = self._gate([B], self.rotate90 * self.p, i, torch.zeros_like(i), dev)
            if dbg is not None:
                i = torch.full_like(i, torch.floor(dbg * 4))
            G_inv = G_inv @ rotate2d_inv(-np.pi / 2 * i)
        if geom_on and self.xint > 0:
            t = (torch.rand([B, 2], device=dev) * 2 - 1) * self.xint_max
            t = self._gate([B, 1], self.xint * self.p, t, torch.zeros_like(t), dev)
            if dbg is not None:
                t = torch.full_like(t, (dbg * 2 - 1) * self.xint_max)
            G_inv = G_inv @ translate2d_inv(torch.round(t[:, 0] * W), torch.round(t[:, 1] * H))
        if geom_on and self.scale > 0:
            s = torch.exp2(torch.randn([B], device=dev) * self.scale_std)
            s = self._gate([B], self.scale * self.p, s, torch.ones_like(s), dev)
            if dbg is not None:
                s = torch.full_like(s, torch.exp2(torch.erfinv(dbg * 2 - 1) * self.scale_std))
            G_inv = G_inv @ scale2d_inv(s, s)
        p_rot = 1 - torch.sqrt((1 - self.rotate * self.p).clamp(0, 1)) if geom_on else None      # P(pre or post) = rotate * p
        if geom_on and self.rotate > 0:
            th = (torch.rand([B], device=dev) * 2 - 1) * np.pi * self.rotate_max
            th = self._gate([B], p_rot, th, torch.zeros_like(th), dev)
            if dbg is not None:
                th = torch.full_like(th, (dbg * 2 - 1) * np.pi * self.rotate_max)
            G_inv = G_inv @ rotate2d_inv(-th)
        if geom_on and self.aniso > 0:
            s = torch.exp2(torch.randn([B], device=dev) * self.aniso_std)
            s = self._gate([B], self.aniso * self.p, s, torch.ones_like(s), dev)
            if dbg is not None:
                s = torch.full_like(s, torch.exp2(torch.erfinv(dbg * 2 - 1) * self.aniso_std))
            G_inv = G_inv @ scale2d_inv(s, 1 / s)
        if geom_on and self.rotate > 0:
            th = (torch.rand([B], device=dev) * 2 - 1) * np.pi * self.rotate_max
            th = self._gate([B], p_rot, th, torch.zeros_like(th), dev)
            if dbg is not None:
                th = torch.zeros_like(th)
            G_inv = G_inv @ rotate2d_inv(-th)
        if geom_on and self.xfrac > 0:
            t = torch.randn([B, 2], device=dev) * self.xfrac_std
            t = self._gate([B, 1], self.xfrac * self.p, t, torch.zeros_like(t), dev)
            if dbg is not None:
                t = torch.full_like(t, torch.erfinv(dbg * 2 - 1) * self.xfrac_std)
            G_inv = G_inv @ translate2d_inv(t[:, 0] * W, t[:, 1] * H)

        # ---------------- execute the warp ----------------
        if G_inv is not I_3:
            images = self._warp(images, G_inv)

        # ---------------- colour parameters -> C ----------------
        I_4 = torch.eye(4, device=dev)
        Cm = I_4
        if self.brightness > 0:
            b = torch.randn([B], device=dev) * self.brightness_std
            b = self._gate([B], self.brightness * self.p, b, torch.zeros_like(b), dev)
            if dbg is not None:
                b = torch.full_like(b, torch.erfinv(dbg * 2 - 1) * self.brightness_std)
            Cm = translate3d(b, b, b) @ Cm
        if self.contrast > 0:
            c = torch.exp2(torch.randn([B], device=dev) * self.contrast_std)
            c = self._gate([B], self.contrast * self.p, c, torch.ones_like(c), dev)
            if dbg is not None:
                c = torch.full_like(c, torch.exp2(torch.erfinv(dbg * 2 - 1) * self.contrast_std))
            Cm = scale3d(c, c, c) @ Cm
        v = misc.constant(np.asarray([1, 1, 1, 0]) / np.sqrt(3), device=dev)    # luma axis
        if self.lumaflip > 0:
            i = torch.floor(torch.rand([B, 1, 1], device=dev) * 2)
            i = self._gate([B, 1, 1], self.lumaflip * self.p, i, torch.zeros_like(i), dev)
            if dbg is not None:
                i = torch.full_like(i, torch.floor(dbg * 2))
            Cm = (I_4 - 2 * v.ger(v) * i) @ Cm                                   # Householder reflection about luma
        if self.hue > 0 and C > 1:
            th = (torch.rand([B], device=dev) * 2 - 1) * np.pi * self.hue_max
            th = self._gate([B], self.hue * self.p, th, torch.zeros_like(th), dev)
            if dbg is not None:
                th = torch.full_like(th, (dbg * 2 - 1) * np.pi * self.hue_max)
            Cm = rotate3d(v, th) @ Cm
        if self.saturation > 0 and C > 1:
            s = torch.exp2(torch.randn([B, 1, 1], device=dev) * self.saturation_std)
            s = self._gate([B, 1, 1], self.saturation * self.p, s, torch.ones_like(s), dev)
            if dbg is not None:
                s = torch.full_like(s, torch.exp2(torch.erfinv(dbg * 2 - 1) * self.saturation_std))
            Cm = (v.ger(v) + (I_4 - v.ger(v)) * s) @ Cm

        if Cm is not I_4:
            images = images.reshape([B, C, H * W])
            if C == 3:
                images = Cm[:, :3, :3] @ images + Cm[:, :3, 3:]
            elif C == 1:
                row = Cm[:, :3, :].mean(dim=1, keepdims=True)
                images = images * row[:, :, :3].sum(dim=2, keepdims=True) + row[:, :, 3:]
            else:
                raise ValueError('Image must be RGB (3 channels) or L (1 channel)')
            images = images.reshape([B, C, H, W])

        # ---------------- image-space band filter ----------------
        if self.imgfilter > 0:
            nb = self.Hz_fbank.shape[0]
            assert len(self.imgfilter_bands) == nb
            expected_power = misc.constant(np.array([10, 1, 1, 1]) / 13, device=dev)      # 1/f spectrum
            g = torch.ones([B, nb], device=dev)
            for i, strength in enumerate(self.imgfilter_bands):
                t_i = torch.exp2(torch.randn([B], device=dev) * self.imgfilter_std)
                t_i = self._gate([B], self.imgfilter * self.p * strength, t_i, torch.ones_like(t_i), dev)
                if dbg is not None:
                    t_i = torch.full_like(t_i, torch.exp2(torch.erfinv(dbg * 2 - 1) * self.imgfilter_std)) if strength > 0 else torch.ones_like(t_i)
                t = torch.ones([B, nb], device=dev)
                t[:, i] = t_i
                t = t / (expected_power * t.square()).sum(dim=-1, keepdims=True).sqrt()
                g = g * t
            taps = (g @ self.Hz_fbank).unsqueeze(1).repeat([1, C, 1]).reshape([B * C, 1, -1])
            pad = self.Hz_fbank.shape[1] // 2
            images = images.reshape([1, B * C, H, W])
            images = torch.nn.functional.pad(input=images, pad=[pad, pad, pad, pad], mode='reflect')
            images = conv2d_gradfix.conv2d(input=images, weight=taps.unsqueeze(2), groups=B * C)
            images = conv2d_gradfix.conv2d(input=images, weight=taps.unsqueeze(3), groups=B * C)
            images = images.reshape([B, C, H, W])

        # ---------------- corruptions ----------------
        if self.noise > 0:
            sigma = torch.randn([B, 1, 1, 1], device=dev).abs() * self.noise_std
            sigma = self._gate([B, 1, 1, 1], self.noise * self.p, sigma, torch.zeros_like(sigma), dev)
            if dbg is not None:
                sigma = torch.full_like(sigma, torch.erfinv(dbg) * self.noise_std)
            images = images + torch.randn([B, C, H, W], device=dev) * sigma
        if self.cutout > 0:
            size = torch.full([B, 2, 1, 1, 1], self.cutout_size, device=dev)
            size = self._gate([B, 1, 1, 1, 1], self.cutout * self.p, size, torch.zeros_like(size), dev)
            center = torch.rand([B, 2, 1, 1, 1], device=dev)
            if dbg is not None:
                size = torch.full_like(size, self.cutout_size)
                center = torch.full_like(center, dbg)
            cx = torch.arange(W, device=dev).reshape([1, 1, 1, -1])
            cy = torch.arange(H, device=dev).reshape([1, 1, -1, 1])
            keep_x = (((cx + 0.5) / W - center[:, 0]).abs() >= size[:, 0] / 2)
            keep_y = (((cy + 0.5) / H - center[:, 1]).abs() >= size[:, 1] / 2)
            images = images * torch.logical_or(keep_x, keep_y).to(torch.float32)
        return images

    def _warp_fused_params(self, images):
        """Geometric augmentation with the parameter algebra in one kernel: the draws below are the reference's draws, in its order
        and shapes (S3/training/augment_mi.py:213-276), so a fixed seed gives the reference's transform parameters."""
        B, C, H, W = images.shape
        dev = images.device
        draws = []

        def draw(on, value_shape, gate_shape, normal):
            if not on:
                draws.extend([None, None])
                return
            v = torch.randn(value_shape, device=dev) if normal else torch.rand(value_shape, device=dev)
            g = torch.rand(gate_shape, device=dev)
            draws.extend([v, g])
        draw(self.xflip > 0, [B], [B], False)
        draw(self.rotate90 > 0, [B], [B], False)
        draw(self.xint > 0, [B, 2], [B, 1], False)
        draw(self.scale > 0, [B], [B], True)
        draw(self.rotate > 0, [B], [B], False)
        draw(self.aniso > 0, [B], [B], True)
        draw(self.rotate > 0, [B], [B], False)
        draw(self.xfrac > 0, [B, 2], [B, 1], True)
        Hz_pad = self.Hz_geom.shape[0] // 4
        theta, margins = aug_warp.params(draws, self.p, (self.xflip, self.rotate90, self.xint, self.scale, self.rotate, self.aniso, self.xfrac),
                                         (self.xint_max, self.scale_std, self.rotate_max, self.aniso_std, self.xfrac_std), B, H, W, Hz_pad)
        out_hw = [(H + Hz_pad * 2) * 2, (W + Hz_pad * 2) * 2]
        images = aug_warp.warp(images, theta, margins, self._hz_geom_taps, out_hw)
        return upfirdn2d.downsample2d(x=images, f=self.Hz_geom, down=2, padding=-Hz_pad * 2, flip_filter=True)

    def _warp(self, images, G_inv):
        """Apply the inverse warp with 2x supersampling (reference :286-321)."""
        B, C, H, W = images.shape
        dev = images.device
        # Margin: how far the warped output corners reach outside the input, per side, maximised over the batch.
        cx, cy = (W - 1) / 2, (H - 1) / 2
        corners = _mat([[-cx, -cy, 1], [cx, -cy, 1], [cx, cy, 1], [-cx, cy, 1]], dev)           # [4, xyz]
        cp = G_inv @ corners.t()                                                                # [B, xyz, 4]
        Hz_pad = self.Hz_geom.shape[0] // 4
        m = cp[:, :2, :].permute(1, 0, 2).flatten(1)                                            # [xy, B*4]
        m = torch.cat([-m, m]).max(dim=1).values                                                # [x0, y0, x1, y1]
        m = m + misc.constant([Hz_pad * 2 - cx, Hz_pad * 2 - cy] * 2, device=dev)
        m = m.max(misc.constant([0, 0] * 2, device=dev))
        m = m.min(misc.constant([W - 1, H - 1] * 2, device=dev))
        if self.fused_warp and images.is_cuda and images.dtype == torch.float32:
            # Margins stay on the device; pad + upsample + affine resampling run as one gather kernel.
            mf = m.ceil()
            G_inv = translate2d((mf[0] - mf[2]) / 2, (mf[1] - mf[3]) / 2, device=dev) @ G_inv
            G_inv = scale2d(2, 2, device=dev) @ G_inv @ scale2d_inv(2, 2, device=dev)
            G_inv = translate2d(-0.5, -0.5, device=dev) @ G_inv @ translate2d_inv(-0.5, -0.5, device=dev)
            wu = (mf[0] + mf[2] + W) * 2
            hu = (mf[1] + mf[3] + H) * 2
            shape = [B, C, (H + Hz_pad * 2) * 2, (W + Hz_pad * 2) * 2]
            G_inv = scale2d(2 / wu, 2 / hu, device=dev) @ G_inv @ scale2d_inv(2 / shape[3], 2 / shape[2], device=dev)
            theta = G_inv[:, :2, :].to(torch.float32).contiguous()
            images = aug_warp.warp(images, theta, mf.to(torch.int32).contiguous(), self._hz_geom_taps, shape[2:])
            return upfirdn2d.downsample2d(x=images, f=self.Hz_geom, down=2, padding=-Hz_pad * 2, flip_filter=True)

        mx0, my0, mx1, my1 = (int(v) for v in m.ceil().to(torch.int32).tolist())                # one device->host sync

        images = torch.nn.functional.pad(input=images, pad=[mx0, mx1, my0, my1], mode='reflect')
        G_inv = translate2d((mx0 - mx1) / 2, (my0 - my1) / 2, device=dev) @ G_inv

        images = upfirdn2d.upsample2d(x=images, f=self.Hz_geom, up=2)
        G_inv = scale2d(2, 2, device=dev) @ G_inv @ scale2d_inv(2, 2, device=dev)
        G_inv = translate2d(-0.5, -0.5, device=dev) @ G_inv @ translate2d_inv(-0.5, -0.5, device=dev)

        shape = [B, C, (H + Hz_pad * 2) * 2, (W + Hz_pad * 2) * 2]
        G_inv = scale2d(2 / images.shape[3], 2 / images.shape[2], device=dev) @ G_inv @ scale2d_inv(2 / shape[3], 2 / shape[2], device=dev)
        grid = torch.nn.functional.affine_grid(theta=G_inv[:, :2, :], size=shape, align_corners=False)
        images = grid_sample_gradfix.grid_sample(images, grid)

        return upfirdn2d.downsample2d(x=images, f=self.Hz_geom, down=2, padding=-Hz_pad * 2, flip_filter=True)
