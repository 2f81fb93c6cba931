"""One launch of the CTA-pair halo convolution per shape (for ncu).  python tools/run_pair_conv.py [variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_track_b200 import _lib  # noqa: E402
from gan_track_b200.torch_utils.ops import conv_igemm  # noqa: E402

v = int(sys.argv[1]) if len(sys.argv) > 1 else 5
_lib.load().gt_conv_igemm_config(v)
dev = torch.device('cuda', 0)
for ci, co, r in [(512, 512, 32), (64, 64, 256)]:
    x = torch.randn([32, ci, r, r], device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn([co, ci, 3, 3], device=dev) / (ci * 9) ** 0.5).to(torch.float16)
    pk = conv_igemm.pack_weight(w, False)
    for _ in range(2):
        conv_igemm.igemm_forward(x, w, transpose=False, output_padding=(0, 0), stride=(1, 1), padding=(1, 1), groups=1, packed=pk)
torch.cuda.synchronize()
for ci, co, r in [(512, 512, 32), (128, 128, 128), (64, 64, 256)]:
    x = torch.randn([32, ci, r, r], device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn([co, ci, 3, 3], device=dev) / (ci * 9) ** 0.5).to(torch.float16)
    pk = conv_igemm.pack_weight(w, False)
    for vv in (0, 6):
        _lib.load().gt_conv_igemm_config(vv)
        f = lambda: conv_igemm.igemm_forward(x, w, transpose=False, output_padding=(0, 0), stride=(1, 1), padding=(1, 1), groups=1, packed=pk)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            f()
        e.record(); e.synchronize()
        print(f'{ci}->{co} @{r}: variant {vv} {s.elapsed_time(e) / 20 * 1e3:.1f} us', flush=True)
print('ok')
