"""GPU check of the 3 x TF32 fp32 convolution route against a float64 reference.  python tools/test_tf32x3.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from gan_track_b200.torch_utils.ops import conv_igemm  # noqa: E402
from test_gpu_conv_igemm import FP32_CASES  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
for name, N, ci, co, H, W, k, s, p, tr in FP32_CASES:
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn([N, ci, H, W], device='cuda', generator=g)
    wshape = [ci, co, k, k] if tr else [co, ci, k, k]
    w = torch.randn(wshape, device='cuda', generator=g) / (ci * k * k) ** 0.5
    y = conv_igemm.igemm_forward(x, w, transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.double(), w.double(), stride=s, padding=p) if tr else F.conv2d(x.double(), w.double(), stride=s, padding=p)
    lib32 = F.conv_transpose2d(x, w, stride=s, padding=p) if tr else F.conv2d(x, w, stride=s, padding=p)
    e = float((y.double() - ref).abs().max() / ref.abs().max())
    e32 = float((lib32.double() - ref).abs().max() / ref.abs().max())
    print(f'{name:32s} ours rel err {e:.3e}   cuDNN fp32 rel err {e32:.3e}', flush=True)
