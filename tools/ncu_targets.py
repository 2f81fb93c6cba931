"""The hot kernels at their training shapes (batch 32, 256x256 config), one launch each after a warm-up pass -- the
command the `ncu --set full` captures under profiles/ are taken on.  Usage: python tools/ncu_targets.py [--reps 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_track_b200.torch_utils.ops import aug_warp, bias_act, conv_igemm, modulated, upfirdn2d  # noqa: E402
from gan_track_b200.training import augment  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--reps', type=int, default=1)
a = ap.parse_args()
dev = torch.device('cuda', 0)
torch.manual_seed(0)
cl = torch.channels_last


def t(shape):
    return torch.randn(shape, device=dev).to(torch.float16).contiguous(memory_format=cl)


cfg = dict(output_padding=(0, 0), groups=1)
x512 = t([32, 512, 32, 32]); w512 = t([512, 512, 3, 3]) * 0.015           # G b32 conv1
x256 = t([32, 256, 64, 64]); w256 = t([256, 256, 3, 3]) * 0.02            # G b64 conv1 / D b64 conv0
x64 = t([32, 64, 256, 256]); w64 = t([64, 64, 3, 3]) * 0.04               # G b256 conv1 / D b256 conv0
x128 = t([32, 128, 128, 128]); wT = t([128, 64, 3, 3]) * 0.03             # G b256 conv0 (transposed stride 2)
w128 = t([128, 128, 3, 3]) * 0.03                                           # G b128 conv1 / D b128 conv0
xd = t([32, 64, 257, 257]); wd = t([128, 64, 3, 3]) * 0.04                  # D b256 conv1 (blurred, stride 2)
x32f = torch.randn([32, 512, 16, 16], device=dev); w32f = torch.randn([512, 512, 3, 3], device=dev) * 0.015   # G b16 conv1 (fp32)
xb = t([32, 64, 257, 257])
xu = t([32, 64, 128, 128])                                                # backward of the D skip downsample
f = upfirdn2d.setup_filter([1, 3, 3, 1], device=dev)
b = torch.randn([64], device=dev, dtype=torch.float16)
pipe = augment.AugmentPipe(xflip=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1)
img = torch.randn([32, 1, 256, 256], device=dev)
th = torch.tensor([[0.97, 0.02, 0.01], [-0.02, 0.98, -0.01]], device=dev).repeat(32, 1, 1).contiguous()
mg = torch.tensor([9, 9, 9, 9], device=dev, dtype=torch.int32)
xp32 = torch.randn([32, 512, 33, 33], device=dev)
sm = torch.randn([32, 64], device=dev)
dm = torch.rand([32, 64], device=dev)
nz = torch.randn([32, 1, 256, 256], device=dev).to(torch.float16)

for _rep in range(a.reps + 1):
    if _rep == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()      # ncu --profile-from-start off: capture the passes after the warm-up pass only
    conv_igemm.igemm_forward(x512, w512, transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    conv_igemm.igemm_forward(x256, w256, transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    conv_igemm.igemm_forward(x64, w64, transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    conv_igemm.igemm_forward(x128, wT, transpose=True, stride=(2, 2), padding=(0, 0), **cfg)
    conv_igemm.igemm_forward(x128, w128, transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    conv_igemm.igemm_forward(xd, wd, transpose=False, stride=(2, 2), padding=(0, 0), **cfg)
    conv_igemm.igemm_wgrad(x512, x512, (512, 512, 3, 3), transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    conv_igemm.igemm_wgrad(x256, x256, (256, 256, 3, 3), transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    conv_igemm.igemm_wgrad(x128, x128, (128, 128, 3, 3), transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    conv_igemm.igemm_wgrad(x64, x64, (64, 64, 3, 3), transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    conv_igemm.igemm_forward(x32f, w32f, transpose=False, stride=(1, 1), padding=(1, 1), **cfg)     # fp32 block on the fp16 x 3 route
    conv_igemm.igemm_wgrad(x32f, x32f, (512, 512, 3, 3), transpose=False, stride=(1, 1), padding=(1, 1), **cfg)
    upfirdn2d.upfirdn2d(xb, f, padding=[1, 1, 1, 1], gain=4)
    upfirdn2d.upfirdn2d(x64, f, down=2, padding=[1, 1, 1, 1])
    upfirdn2d.upfirdn2d(xu, f, up=2, padding=[2, 1, 2, 1], gain=4)
    upfirdn2d.upfirdn2d(xp32, f, padding=[1, 1, 1, 1], gain=4)
    xg = x64.clone().requires_grad_(True)
    bg = b.clone().requires_grad_(True)
    y = bias_act.bias_act(xg, bg, act='lrelu', clamp=256.0)
    torch.autograd.grad(y, [xg, bg], x64)
    sg, dg, ng = sm.clone().requires_grad_(True), dm.clone().requires_grad_(True), nz.clone().requires_grad_(True)
    ym = modulated.mod_scale(xg, sg)
    torch.autograd.grad(ym, [xg, sg], x64)
    yd = modulated.demod_act(xg, dg, ng, bg, act='lrelu', gain=1.4142, clamp=256.0)
    torch.autograd.grad(yd, [xg, dg, ng, bg], x64)
    aug_warp.warp(img, th, mg, pipe._hz_geom_taps, (524, 524))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('ok')
