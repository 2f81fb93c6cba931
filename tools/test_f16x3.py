"""fp32 convolutions of the low-resolution blocks: the fp16 x 3 tensor-core route (csrc/conv_f16x3.cu) against the library's fp32 kernels
(TF32 off) -- accuracy vs float64 and time per call (forward, data gradient, weight gradient) at batch 32.  python tools/test_f16x3.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    from gan_track_b200.torch_utils.ops import conv_igemm
    dev = torch.device('cuda', 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    flush = torch.empty([256 << 20], dtype=torch.uint8, device=dev)

    def bench(fn, iters=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            fn()
            e_.record()
            e_.synchronize()
            ts.append(s_.elapsed_time(e_))
        return sum(ts) / len(ts) * 1e3

    def rel(a, b):
        return float((a.double() - b).abs().max() / b.abs().max())
    cases = [('G b4 conv1 512->512 3x3 @4', 32, 512, 512, 4, 3, 1, 1, False), ('G b8 conv0 512->512 T s2 @4->9', 32, 512, 512, 4, 3, 2, 0, True),
             ('G b8 conv1 512->512 3x3 @8', 32, 512, 512, 8, 3, 1, 1, False), ('G b16 conv0 512->512 T s2 @8->17', 32, 512, 512, 8, 3, 2, 0, True),
             ('G b16 conv1 512->512 3x3 @16', 32, 512, 512, 16, 3, 1, 1, False), ('D b16 conv1 512->512 s2 @17', 32, 512, 512, 17, 3, 2, 0, False),
             ('D b16 skip 512->512 1x1 @8', 32, 512, 512, 8, 1, 1, 0, False), ('D b4 conv 513->512 3x3 @4', 32, 513, 512, 4, 3, 1, 1, False),
             ('D b16 conv0 512->512 3x3 @16 n64', 64, 512, 512, 16, 3, 1, 1, False)]
    print(f'# {"case":36s} {"fwd ours":>9s} {"lib":>8s} {"err ours":>9s} {"err lib":>8s} | {"dgrad ours":>10s} {"lib":>8s} | {"wgrad ours":>10s} {"lib":>8s} {"err ours":>9s} {"err lib":>8s}   (us; errors vs float64)')
    for name, N, ci, co, R, k, s, p, tr in cases:
        x = torch.randn([N, ci, R, R], device=dev).contiguous(memory_format=torch.channels_last)
        wshape = [ci, co, k, k] if tr else [co, ci, k, k]
        w = torch.randn(wshape, device=dev) / (ci * k * k) ** 0.5
        kw = dict(transpose=tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
        lib = (lambda: F.conv_transpose2d(x, w, stride=s, padding=p)) if tr else (lambda: F.conv2d(x, w, stride=s, padding=p))
        y = conv_igemm.igemm_forward(x, w, **kw)
        yl = lib()
        ref = (F.conv_transpose2d(x.double(), w.double(), stride=s, padding=p) if tr else F.conv2d(x.double(), w.double(), stride=s, padding=p))
        t_f, t_fl = bench(lambda: conv_igemm.igemm_forward(x, w, **kw)), bench(lib)
        dy = torch.randn_like(yl).contiguous(memory_format=torch.channels_last)
        kwd = dict(transpose=not tr, output_padding=(0, 0), stride=(s, s), padding=(p, p), groups=1)
        t_d = bench(lambda: conv_igemm.igemm_forward(dy, w, **kwd))
        t_dl = bench(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [s, s], [p, p], [1, 1], tr, [0, 0], 1, [True, False, False]))
        dw = conv_igemm.igemm_wgrad(dy, x, tuple(wshape), **kw)
        dwl = torch.ops.aten.convolution_backward(dy, x, w, None, [s, s], [p, p], [1, 1], tr, [0, 0], 1, [False, True, False])[1]
        dwr = torch.ops.aten.convolution_backward(dy.double(), x.double(), w.double(), None, [s, s], [p, p], [1, 1], tr, [0, 0], 1, [False, True, False])[1]
        t_w = bench(lambda: conv_igemm.igemm_wgrad(dy, x, tuple(wshape), **kw))
        t_wl = bench(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [s, s], [p, p], [1, 1], tr, [0, 0], 1, [False, True, False]))
        print(f'  {name:36s} {t_f:9.1f} {t_fl:8.1f} {rel(y, ref):9.1e} {rel(yl, ref):8.1e} | {t_d:10.1f} {t_dl:8.1f} | {t_w:10.1f} {t_wl:8.1f} {rel(dw, dwr):9.1e} {rel(dwl, dwr):8.1e}',
              flush=True)


if __name__ == '__main__':
    main()
