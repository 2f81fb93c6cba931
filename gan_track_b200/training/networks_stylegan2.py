"""StyleGAN2 generator and discriminator: the CALLERS of the op surface.

Same classes, constructor arguments, parameter / buffer names and forward semantics as the reference
(S3/training/networks_stylegan2.py: modulated_conv2d :32-89, FullyConnectedLayer :94, Conv2dLayer :133, MappingNetwork
:191, SynthesisLayer :274, ToRGBLayer :338, SynthesisBlock :363, SynthesisNetwork :467, Generator :527,
DiscriminatorBlock :555, MinibatchStdLayer :644, DiscriminatorEpilogue :673, Discriminator :734), so a reference
state_dict loads unchanged and `tests/` can compare outputs tensor by tensor.  The host code only fixes shapes, dtypes,
memory formats, gains and clamps; all arithmetic goes through `torch_utils.ops`.

One deliberate default differs: fp16 blocks use channels-last memory (`fp16_channels_last=True`), the layout the
tcgen05 implicit-GEMM kernels consume directly (DESIGN.md "data layout").  The reference exposes the same switch
(:375, :568) with default False; results are layout-independent.
"""
import numpy as np
import torch

from ..torch_utils import misc
from ..torch_utils.ops import bias_act, conv2d_resample, fc, fma, modulated, rgb, upfirdn2d

SQRT_HALF = np.sqrt(0.5)
ARCHITECTURES = ('orig', 'skip', 'resnet')


# ---------------------------------------------------------------------------------------------------------------------
# small shared pieces
# ---------------------------------------------------------------------------------------------------------------------

def _keep(module, **attrs):
    """Plain (non-parameter) attributes of a layer, set in one place."""
    for key, value in attrs.items():
        setattr(module, key, value)


def _layout(channels_last):
    return torch.channels_last if channels_last else torch.contiguous_format


# fp32 blocks on a CUDA device also compute channels-last: their convolutions run on the NHWC tensor-core kernels (fp16 x 3 route,
# ops/conv_igemm.py) and the fused modulation / activation kernels take dense NHWC tensors.  Results are layout independent.
fp32_channels_last = True


def _working_type(use_fp16, channels_last, force_fp32, on_cuda=True):
    """(dtype, memory_format) a block computes in: fp16 (+ channels-last) unless forced to fp32 or off the GPU."""
    half = use_fp16 and not force_fp32 and on_cuda
    if not on_cuda:
        return torch.float32, torch.contiguous_format
    if half:
        return torch.float16, _layout(channels_last)
    return torch.float32, _layout(fp32_channels_last)


def _conv_weight(out_channels, in_channels, kernel_size, channels_last):
    return torch.randn([out_channels, in_channels, kernel_size, kernel_size]).to(memory_format=_layout(channels_last))


def _scaled(clamp, gain):
    return None if clamp is None else clamp * gain


def _describe(module, *names):
    return ', '.join(f'{n}={getattr(module, n)}' for n in names)


def _fp16_from(resolution_log2, num_fp16_res):
    """Lowest resolution that runs in fp16: the top `num_fp16_res` resolutions, never below 8."""
    return max(2 ** (resolution_log2 + 1 - num_fp16_res), 8)


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


# ---------------------------------------------------------------------------------------------------------------------
# modulated convolution
# ---------------------------------------------------------------------------------------------------------------------

# fused_modconv=True (eval-mode sampling: per-sample weights W * s * d and one grouped convolution, reference :79-89) is
# evaluated on CUDA tensors as scale -> shared-weight convolution -> demodulate, which is the same function (4e-7 in fp32,
# SURVEY A.3; pinned against the reference's own eval-mode output in tests/test_gpu_model.py) and keeps G_ema sampling on the
# tcgen05 / fused element-wise kernels: 24.8 -> 3.9 ms per batch of 32 at 256x256 (tools/bench_sampling.py).  True restores the grouped convolution.
grouped_fused_modconv = False


def _demod_operands(x, weight, styles, shared_weight_route):
    """(weight, styles, dcoefs) for a demodulated layer.  In fp16 both operands are pre-normalised by their infinity norms so
    that neither the scaled activations nor the convolution overflow (demodulation makes the result invariant to it);
    dcoefs[n,o] = rsqrt(sum_i s[n,i]^2 * sum_k W[o,i,k]^2 + eps), i.e. the reduction of the reference's [N,O,I,kh,kw] product
    (:60-63) without materialising it."""
    half = x.dtype == torch.float16
    if shared_weight_route and modulated.prep_applicable(weight, styles):
        # pre-normalisation and demodulation coefficients in four launches (csrc/modprep.cu) instead of ~13 tensor ops
        w16, sn, dcoefs = modulated.prep(weight, styles, half)
        return (w16, sn, dcoefs) if half else (weight, styles, dcoefs)
    if half:
        fan_in = weight.shape[1] * weight.shape[2] * weight.shape[3]
        weight = weight * (1 / np.sqrt(fan_in) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    per_pair_energy = weight.square().sum(dim=[2, 3])                          # [O, I]
    return weight, styles, (styles.square() @ per_pair_energy.t() + 1e-8).rsqrt()   # dcoefs [N, O]


def _is_pixel_dot(x, weight, up, down, padding):
    """ToRGB to one image channel on a channels-last CUDA tensor: a per-pixel dot product."""
    out_channels, _, kh, kw = weight.shape
    return (out_channels == 1 and kh == 1 and kw == 1 and up == 1 and down == 1 and padding == 0 and x.is_cuda and x.dim() == 4
            and x.stride(1) == 1 and x.is_contiguous(memory_format=torch.channels_last))


def modulated_conv2d(
    x,                          # [N, Cin, H, W]
    weight,                     # [Cout, Cin, kh, kw]
    styles,                     # [N, Cin] modulation coefficients
    noise=None,                 # broadcastable to the output, added after demodulation
    up=1, down=1, padding=0,
    resample_filter=None,       # from upfirdn2d.setup_filter()
    demodulate=True,
    flip_weight=True,           # True = correlation (torch conv2d)
    fused_modconv=True,         # True: per-sample weights + grouped conv; False: scale activations around a shared-weight conv
    bias_act_args=None,         # (not in the reference) dict(b, act, gain, clamp): apply the layer's bias_act here, so that on
                                # channels-last CUDA tensors demodulation + noise + bias + activation run as one fused pass
):
    n = x.shape[0]
    out_channels, in_channels, kh, kw = weight.shape
    misc.assert_shape(x, [n, in_channels, None, None])
    misc.assert_shape(styles, [n, in_channels])
    grouped = fused_modconv and not (x.is_cuda and not grouped_fused_modconv)
    resample = dict(f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
    dcoefs = None
    if demodulate:
        weight, styles, dcoefs = _demod_operands(x, weight, styles, shared_weight_route=not grouped)

    def epilogue(y):
        if bias_act_args is None:
            return y
        return bias_act.bias_act(y, bias_act_args['b'], act=bias_act_args['act'], gain=bias_act_args['gain'], clamp=bias_act_args['clamp'])

    if grouped:
        # per-sample weights W * s (* d), one grouped convolution over the batch folded into the channel axis
        per_sample = weight.unsqueeze(0) * styles.reshape(n, 1, -1, 1, 1)      # [N, O, I, kh, kw]
        if demodulate:
            per_sample = per_sample * dcoefs.reshape(n, -1, 1, 1, 1)
        y = conv2d_resample.conv2d_resample(x=x.reshape(1, -1, *x.shape[2:]), w=per_sample.reshape(-1, in_channels, kh, kw).to(x.dtype), groups=n, **resample)
        y = y.reshape(n, -1, *y.shape[2:])
        return epilogue(y if noise is None else y.add_(noise))

    # shared weight: scale the activations by the styles, convolve once, demodulate the result
    fused_elementwise = modulated.applicable(x)
    x = modulated.mod_scale(x, styles) if fused_elementwise else x * styles.to(x.dtype).reshape(n, -1, 1, 1)
    if _is_pixel_dot(x, weight, up, down, padding):
        # [N,H,W,C] @ [C,1] on a free view.  Same rounding points as the 1x1 convolution (fp32 accumulation, one rounding of the
        # result) and differentiable to any order by autograd -- the form the path-length pass uses; the library's convolution
        # for Cout = 1 costs two fp32 kernels of 470 us plus format copies there.
        y = torch.matmul(x.permute(0, 2, 3, 1), weight.to(x.dtype).reshape(in_channels, 1)).permute(0, 3, 1, 2)
    else:
        y = conv2d_resample.conv2d_resample(x=x, w=weight.to(x.dtype), **resample)
    if bias_act_args is not None and bias_act_args['act'] in ('linear', 'lrelu') and modulated.applicable(y):
        spec = bias_act.activation_funcs[bias_act_args['act']]
        gain = spec.def_gain if bias_act_args['gain'] is None else bias_act_args['gain']
        return modulated.demod_act(y, dcoefs, noise, bias_act_args['b'], act=bias_act_args['act'], alpha=spec.def_alpha, gain=gain,
                                   clamp=bias_act_args['clamp'])
    if dcoefs is not None:
        d = dcoefs.to(y.dtype).reshape(n, -1, 1, 1)
        y = y * d if noise is None else fma.fma(y, d, noise.to(y.dtype))
    elif noise is not None:
        y = y.add_(noise.to(y.dtype))
    return epilogue(y)


# ---------------------------------------------------------------------------------------------------------------------
# layers
# ---------------------------------------------------------------------------------------------------------------------

class FullyConnectedLayer(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        _keep(self, in_features=in_features, out_features=out_features, activation=activation,
              weight_gain=lr_multiplier / np.sqrt(in_features), bias_gain=lr_multiplier)
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None

    def forward(self, x):
        linear = self.activation == 'linear'
        if fc.applicable(x, self.weight):
            # one kernel for gains + GEMM + bias (csrc/fc.cu); the activation of non-linear layers stays with bias_act
            y = fc.linear(x, self.weight, self.bias, self.weight_gain, self.bias_gain)
            return y if linear else bias_act.bias_act(y, None, act=self.activation)
        w = self.weight.to(x.dtype) * self.weight_gain
        b = None
        if self.bias is not None:
            b = self.bias.to(x.dtype)
            b = b if self.bias_gain == 1 else b * self.bias_gain
        if linear and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return bias_act.bias_act(x.matmul(w.t()), b, act=self.activation)

    def extra_repr(self):
        return _describe(self, 'in_features', 'out_features', 'activation')


class Conv2dLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='linear', up=1, down=1,
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False, trainable=True):
        super().__init__()
        _keep(self, in_channels=in_channels, out_channels=out_channels, activation=activation, up=up, down=down, conv_clamp=conv_clamp,
              padding=kernel_size // 2, weight_gain=1 / np.sqrt(in_channels * kernel_size * kernel_size),
              act_gain=bias_act.activation_funcs[activation].def_gain)
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        tensors = dict(weight=_conv_weight(out_channels, in_channels, kernel_size, channels_last), bias=torch.zeros([out_channels]) if bias else None)
        for name, value in tensors.items():                  # frozen layers (freeze_layers of the discriminator) keep them as buffers
            if value is None:
                setattr(self, name, None)
            elif trainable:
                setattr(self, name, torch.nn.Parameter(value))
            else:
                self.register_buffer(name, value)

    def forward(self, x, gain=1, addend=None):
        # addend (not in the reference): a residual of the shape of the result, added to it (DiscriminatorBlock's shortcut.add_(x), :636)
        w = (self.weight * self.weight_gain).to(x.dtype)
        b = None if self.bias is None else self.bias.to(x.dtype)
        act = dict(act=self.activation, gain=self.act_gain * gain, clamp=_scaled(self.conv_clamp, gain))
        simple_act = self.activation in ('linear', 'lrelu')
        plain = self.up == 1 and self.down == 1
        if simple_act and plain and self.in_channels == 1 and self.weight.shape[2] == 1 and rgb.applicable(x, self.out_channels):
            # FromRGB on single-channel slices: outer product + bias_act in one pass, written channels-last (csrc/rgb.cu)
            y = rgb.fromrgb1(x, w.reshape(-1), b, **act)
            return y if addend is None else y.add_(addend)
        resample = dict(f=self.resample_filter, up=self.up, down=self.down, padding=self.padding)
        if simple_act and self.up == 1 and x.is_cuda:
            # the convolution is the last kernel of conv2d_resample: its bias_act (and the residual) ride in the convolution's epilogue
            return conv2d_resample.conv2d_resample(x=x, w=w, flip_weight=True, epilogue=dict(b=b, addend=addend, **act), **resample)
        y = conv2d_resample.conv2d_resample(x=x, w=w, flip_weight=(self.up == 1), **resample)
        y = bias_act.bias_act(y, b, **act)
        return y if addend is None else y.add_(addend)

    def extra_repr(self):
        return _describe(self, 'in_channels', 'out_channels', 'activation', 'up', 'down')


class MappingNetwork(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None, activation='lrelu',
                 lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        _keep(self, z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=num_ws, num_layers=num_layers, w_avg_beta=w_avg_beta)
        embed_features = 0 if c_dim == 0 else (w_dim if embed_features is None else embed_features)
        hidden = w_dim if layer_features is None else layer_features
        widths = [z_dim + embed_features, *([hidden] * (num_layers - 1)), w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        for idx, (fan_in, fan_out) in enumerate(zip(widths[:-1], widths[1:])):
            setattr(self, f'fc{idx}', FullyConnectedLayer(fan_in, fan_out, activation=activation, lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, ema_rows=None):
        # ema_rows (not in the reference): only the first `ema_rows` rows feed the w_avg update -- lets a caller map two latent
        # batches in one pass while tracking the average of the first, as two separate calls would
        parts = []
        if self.z_dim > 0:
            misc.assert_shape(z, [None, self.z_dim])
            parts.append(normalize_2nd_moment(z.to(torch.float32)))
        if self.c_dim > 0:
            misc.assert_shape(c, [None, self.c_dim])
            parts.append(normalize_2nd_moment(self.embed(c.to(torch.float32))))
        x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if update_emas and self.w_avg_beta is not None:
            self.w_avg.copy_(x.detach()[:ema_rows].mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x

    def extra_repr(self):
        return _describe(self, 'z_dim', 'c_dim', 'w_dim', 'num_ws')


class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True, activation='lrelu',
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
        super().__init__()
        _keep(self, in_channels=in_channels, out_channels=out_channels, w_dim=w_dim, resolution=resolution, up=up, use_noise=use_noise,
              activation=activation, conv_clamp=conv_clamp, padding=kernel_size // 2, act_gain=bias_act.activation_funcs[activation].def_gain)
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(_conv_weight(out_channels, in_channels, kernel_size, channels_last))
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    def _noise(self, x, noise_mode):
        if not self.use_noise or noise_mode == 'none':
            return None
        if noise_mode == 'const':
            return self.noise_const * self.noise_strength
        return torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1):
        assert noise_mode in ['random', 'const', 'none']
        side = self.resolution // self.up
        misc.assert_shape(x, [None, self.in_channels, side, side])
        styles = self.affine(w)
        noise = self._noise(x, noise_mode)
        act = dict(b=self.bias.to(x.dtype), act=self.activation, gain=self.act_gain * gain, clamp=_scaled(self.conv_clamp, gain))
        return modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=noise, up=self.up, padding=self.padding,
                                resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused_modconv=fused_modconv, bias_act_args=act)

    def extra_repr(self):
        return _describe(self, 'in_channels', 'out_channels', 'w_dim', 'resolution', 'up', 'activation')


class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        _keep(self, in_channels=in_channels, out_channels=out_channels, w_dim=w_dim, conv_clamp=conv_clamp,
              weight_gain=1 / np.sqrt(in_channels * kernel_size * kernel_size))
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(_conv_weight(out_channels, in_channels, kernel_size, channels_last))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, fused_modconv=True):
        styles = self.affine(w) * self.weight_gain
        if not fused_modconv and self.weight.shape[2] == 1 and rgb.torgb_applicable(x, self.out_channels):
            # single-channel slices: modulation + 1x1 convolution + bias + clamp in one pass over x (csrc/rgb.cu)
            return rgb.torgb1(x, self.weight, styles, self.bias, clamp=self.conv_clamp)
        y = modulated_conv2d(x=x, weight=self.weight, styles=styles, demodulate=False, fused_modconv=fused_modconv)
        return bias_act.bias_act(y, self.bias.to(y.dtype), clamp=self.conv_clamp)

    def extra_repr(self):
        return _describe(self, 'in_channels', 'out_channels', 'w_dim')


# ---------------------------------------------------------------------------------------------------------------------
# generator
# ---------------------------------------------------------------------------------------------------------------------

class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=[1, 3, 3, 1], conv_clamp=256, use_fp16=False, fp16_channels_last=True,
                 fused_modconv_default=True, **layer_kwargs):
        assert architecture in ARCHITECTURES
        super().__init__()
        first = in_channels == 0                             # the 4x4 block starts from a learned constant
        has_torgb = is_last or architecture == 'skip'
        _keep(self, in_channels=in_channels, w_dim=w_dim, resolution=resolution, img_channels=img_channels, is_last=is_last,
              architecture=architecture, use_fp16=use_fp16, channels_last=(use_fp16 and fp16_channels_last),
              fused_modconv_default=fused_modconv_default, num_conv=1 if first else 2, num_torgb=int(has_torgb))
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        layer = dict(w_dim=w_dim, conv_clamp=conv_clamp, channels_last=self.channels_last)
        if first:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, resolution=resolution, up=2, resample_filter=resample_filter, **layer, **layer_kwargs)
        self.conv1 = SynthesisLayer(out_channels, out_channels, resolution=resolution, **layer, **layer_kwargs)
        if has_torgb:
            self.torgb = ToRGBLayer(out_channels, img_channels, **layer)
        if not first and architecture == 'resnet':
            self.skip = Conv2dLayer(in_channels, out_channels, kernel_size=1, bias=False, up=2, resample_filter=resample_filter,
                                    channels_last=self.channels_last)

    def forward(self, x, img, ws, force_fp32=False, fused_modconv=None, update_emas=False, **layer_kwargs):
        misc.assert_shape(ws, [None, self.num_conv + self.num_torgb, self.w_dim])
        styles = iter(ws.unbind(dim=1))
        dtype, memory_format = _working_type(self.use_fp16, self.channels_last, force_fp32, ws.device.type == 'cuda')
        fused = self.fused_modconv_default if fused_modconv is None else fused_modconv
        if fused == 'inference_only':
            fused = not self.training
        half_res = self.resolution // 2
        conv = dict(fused_modconv=fused, **layer_kwargs)

        if self.in_channels == 0:
            x = self.const.to(dtype=dtype).unsqueeze(0).repeat([ws.shape[0], 1, 1, 1]).contiguous(memory_format=memory_format)
            x = self.conv1(x, next(styles), **conv)
        else:
            misc.assert_shape(x, [None, self.in_channels, half_res, half_res])
            x = x.to(dtype=dtype, memory_format=memory_format)
            if self.architecture == 'resnet':
                shortcut = self.skip(x, gain=SQRT_HALF)
                x = self.conv0(x, next(styles), **conv)
                x = self.conv1(x, next(styles), gain=SQRT_HALF, **conv)
                x = shortcut.add_(x)
            else:
                x = self.conv0(x, next(styles), **conv)
                x = self.conv1(x, next(styles), **conv)

        if img is not None:
            misc.assert_shape(img, [None, self.img_channels, half_res, half_res])
            img = upfirdn2d.upsample2d(img, self.resample_filter)
        if self.num_torgb:
            rgb_out = self.torgb(x, next(styles), fused_modconv=fused).to(dtype=torch.float32, memory_format=torch.contiguous_format)
            img = rgb_out if img is None else img.add_(rgb_out)

        assert x.dtype == dtype and (img is None or img.dtype == torch.float32)
        return x, img

    def extra_repr(self):
        return _describe(self, 'resolution', 'architecture')


class SynthesisNetwork(torch.nn.Module):
    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=4, **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        log2 = int(np.log2(img_resolution))
        _keep(self, w_dim=w_dim, img_resolution=img_resolution, img_resolution_log2=log2, img_channels=img_channels, num_fp16_res=num_fp16_res,
              block_resolutions=[2 ** i for i in range(2, log2 + 1)], num_ws=0)
        width = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        first_fp16 = _fp16_from(log2, num_fp16_res)
        for res in self.block_resolutions:
            last = res == img_resolution
            block = SynthesisBlock(width[res // 2] if res > 4 else 0, width[res], w_dim=w_dim, resolution=res, img_channels=img_channels,
                                   is_last=last, use_fp16=(res >= first_fp16), **block_kwargs)
            # every conv has its own w; a block's ToRGB shares the next block's first w, so only the last one adds to the count
            self.num_ws += block.num_conv + (block.num_torgb if last else 0)
            setattr(self, f'b{res}', block)

    def forward(self, ws, **block_kwargs):
        misc.assert_shape(ws, [None, self.num_ws, self.w_dim])
        ws = ws.to(torch.float32)
        if ws.is_cuda:
            # one [num_ws, N, w_dim] copy, handed to the blocks as [N, n, w_dim] views whose per-layer slices are contiguous rows:
            # the style affines read them in place instead of 20 per-layer .contiguous() copies (same values)
            ws = ws.transpose(0, 1).contiguous().transpose(0, 1)
        x = img = None
        start = 0
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            x, img = block(x, img, ws.narrow(1, start, block.num_conv + block.num_torgb), **block_kwargs)
            start += block.num_conv
        return img

    def extra_repr(self):
        return _describe(self, 'w_dim', 'num_ws', 'img_resolution', 'img_channels', 'num_fp16_res')


class Generator(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        _keep(self, z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels)
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)


# ---------------------------------------------------------------------------------------------------------------------
# discriminator
# ---------------------------------------------------------------------------------------------------------------------

class DiscriminatorBlock(torch.nn.Module):
    def __init__(self, in_channels, tmp_channels, out_channels, resolution, img_channels, first_layer_idx, architecture='resnet',
                 activation='lrelu', resample_filter=[1, 3, 3, 1], conv_clamp=None, use_fp16=False, fp16_channels_last=True,
                 freeze_layers=0):
        assert in_channels in [0, tmp_channels] and architecture in ARCHITECTURES
        super().__init__()
        _keep(self, in_channels=in_channels, resolution=resolution, img_channels=img_channels, first_layer_idx=first_layer_idx,
              architecture=architecture, use_fp16=use_fp16, channels_last=(use_fp16 and fp16_channels_last), num_layers=0)
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))

        def add(name, cin, cout, kernel_size, **kw):
            """Layers are numbered across the whole discriminator; the first `freeze_layers` of them are not trained."""
            trainable = first_layer_idx + self.num_layers >= freeze_layers
            self.num_layers += 1
            setattr(self, name, Conv2dLayer(cin, cout, kernel_size=kernel_size, trainable=trainable, channels_last=self.channels_last, **kw))

        if in_channels == 0 or architecture == 'skip':
            add('fromrgb', img_channels, tmp_channels, 1, activation=activation, conv_clamp=conv_clamp)
        add('conv0', tmp_channels, tmp_channels, 3, activation=activation, conv_clamp=conv_clamp)
        add('conv1', tmp_channels, out_channels, 3, activation=activation, down=2, resample_filter=resample_filter, conv_clamp=conv_clamp)
        if architecture == 'resnet':
            add('skip', tmp_channels, out_channels, 1, bias=False, down=2, resample_filter=resample_filter)

    def forward(self, x, img, force_fp32=False):
        on_cuda = (img if x is None else x).device.type == 'cuda'
        dtype, memory_format = _working_type(self.use_fp16, self.channels_last, force_fp32, on_cuda)
        square = [self.resolution, self.resolution]
        if x is not None:
            misc.assert_shape(x, [None, self.in_channels, *square])
            x = x.to(dtype=dtype, memory_format=memory_format)
        if self.in_channels == 0 or self.architecture == 'skip':
            misc.assert_shape(img, [None, self.img_channels, *square])
            img = img.to(dtype=dtype, memory_format=memory_format)
            features = self.fromrgb(img)
            x = features if x is None else x + features
            img = upfirdn2d.downsample2d(img, self.resample_filter) if self.architecture == 'skip' else None
        if self.architecture == 'resnet':
            # shortcut.add_(main) of the reference (:636), with the residual handed to the skip layer: on fp16 CUDA tensors the addition
            # runs in the epilogue of the skip convolution (the same fp16 sum of the two rounded branches, one pass over the result less)
            main = self.conv1(self.conv0(x), gain=SQRT_HALF)
            x = self.skip(x, gain=SQRT_HALF, addend=main)
        else:
            x = self.conv1(self.conv0(x))
        assert x.dtype == dtype
        return x, img

    def extra_repr(self):
        return _describe(self, 'resolution', 'architecture')


class MinibatchStdLayer(torch.nn.Module):
    def __init__(self, group_size, num_channels=1):
        super().__init__()
        _keep(self, group_size=group_size, num_channels=num_channels)

    def forward(self, x):
        n, channels, h, w = x.shape
        group = int(n) if self.group_size is None else min(int(self.group_size), int(n))
        stats = self.num_channels
        # [G, n/G, F, C/F, H, W]: member g of group j is sample g * (n/G) + j; channels in F sets
        members = x.reshape(group, -1, stats, channels // stats, h, w)
        spread = (members - members.mean(dim=0)).square().mean(dim=0)       # variance over the group's members
        per_group = (spread + 1e-8).sqrt().mean(dim=[2, 3, 4])               # [n/G, F]
        plane = per_group.reshape(-1, stats, 1, 1).repeat(group, 1, h, w)
        return torch.cat([x, plane], dim=1)

    def extra_repr(self):
        return _describe(self, 'group_size', 'num_channels')


class DiscriminatorEpilogue(torch.nn.Module):
    def __init__(self, in_channels, cmap_dim, resolution, img_channels, architecture='resnet', mbstd_group_size=4,
                 mbstd_num_channels=1, activation='lrelu', conv_clamp=None):
        assert architecture in ARCHITECTURES
        super().__init__()
        _keep(self, in_channels=in_channels, cmap_dim=cmap_dim, resolution=resolution, img_channels=img_channels, architecture=architecture)
        if architecture == 'skip':
            self.fromrgb = Conv2dLayer(img_channels, in_channels, kernel_size=1, activation=activation)
        self.mbstd = MinibatchStdLayer(group_size=mbstd_group_size, num_channels=mbstd_num_channels) if mbstd_num_channels > 0 else None
        self.conv = Conv2dLayer(in_channels + mbstd_num_channels, in_channels, kernel_size=3, activation=activation, conv_clamp=conv_clamp)
        self.fc = FullyConnectedLayer(in_channels * resolution * resolution, in_channels, activation=activation)
        self.out = FullyConnectedLayer(in_channels, cmap_dim if cmap_dim > 0 else 1)

    def forward(self, x, img, cmap, force_fp32=False):
        square = [self.resolution, self.resolution]
        misc.assert_shape(x, [None, self.in_channels, *square])
        fp32 = dict(dtype=torch.float32, memory_format=torch.contiguous_format)
        x = x.to(**fp32)
        if self.architecture == 'skip':
            misc.assert_shape(img, [None, self.img_channels, *square])
            x = x + self.fromrgb(img.to(**fp32))
        if self.mbstd is not None:
            x = self.mbstd(x)
        logits = self.out(self.fc(self.conv(x).flatten(1)))
        if self.cmap_dim > 0:
            # projection discriminator: score = <features, embedding of the class> / sqrt(dim)
            misc.assert_shape(cmap, [None, self.cmap_dim])
            logits = (logits * cmap).sum(dim=1, keepdim=True) * (1 / np.sqrt(self.cmap_dim))
        assert logits.dtype == torch.float32
        return logits

    def extra_repr(self):
        return _describe(self, 'resolution', 'architecture')


class Discriminator(torch.nn.Module):
    def __init__(self, c_dim, img_resolution, img_channels, architecture='resnet', channel_base=32768, channel_max=512,
                 num_fp16_res=4, conv_clamp=256, cmap_dim=None, block_kwargs={}, mapping_kwargs={}, epilogue_kwargs={}):
        super().__init__()
        log2 = int(np.log2(img_resolution))
        _keep(self, c_dim=c_dim, img_resolution=img_resolution, img_resolution_log2=log2, img_channels=img_channels,
              block_resolutions=[2 ** i for i in range(log2, 2, -1)])
        width = {res: min(channel_base // res, channel_max) for res in [*self.block_resolutions, 4]}
        first_fp16 = _fp16_from(log2, num_fp16_res)
        cmap_dim = 0 if c_dim == 0 else (width[4] if cmap_dim is None else cmap_dim)
        shared = dict(img_channels=img_channels, architecture=architecture, conv_clamp=conv_clamp)
        layers_so_far = 0
        for res in self.block_resolutions:
            block = DiscriminatorBlock(width[res] if res < img_resolution else 0, width[res], width[res // 2], resolution=res,
                                       first_layer_idx=layers_so_far, use_fp16=(res >= first_fp16), **block_kwargs, **shared)
            setattr(self, f'b{res}', block)
            layers_so_far += block.num_layers
        if c_dim > 0:
            self.mapping = MappingNetwork(z_dim=0, c_dim=c_dim, w_dim=cmap_dim, num_ws=None, w_avg_beta=None, **mapping_kwargs)
        self.b4 = DiscriminatorEpilogue(width[4], cmap_dim=cmap_dim, resolution=4, **epilogue_kwargs, **shared)

    def forward(self, img, c, update_emas=False, **block_kwargs):
        x = None
        for res in self.block_resolutions:
            x, img = getattr(self, f'b{res}')(x, img, **block_kwargs)
        cmap = self.mapping(None, c) if self.c_dim > 0 else None
        return self.b4(x, img, cmap)

    def extra_repr(self):
        return _describe(self, 'c_dim', 'img_resolution', 'img_channels')
