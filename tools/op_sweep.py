"""Op microbenchmark sweep (BASELINE.json config 5): bias_act / upfirdn2d / one modulated-convolution layer at every resolution
4^2 .. 512^2 of the StyleGAN2 configurations, batch 8 / 32 / 64, fp16 / fp32, this repository's kernels against the reference's
GPU implementation on the same box:

  * bias_act, upfirdn2d: the reference's OWN CUDA plugin (oracle/_ref/*.so, compiled from /root/reference by
    oracle/build_ref_plugin.py), called through the same pybind entry point as ours;
  * modulated_conv2d (training branch, S3/training/networks_stylegan2.py:52-77 + 325-327): the reference's op chain restated
    over its plugin and the library convolution it calls (`x * styles` -> F.conv2d / F.conv_transpose2d -> plugin upfirdn2d ->
    addcmul -> plugin bias_act; NCHW as the reference's default memory format), against `modulated_conv2d` + `bias_act` of this
    package.  Forward only; the per-kernel backward numbers are in bench.py's `roofline_all`.

Timing: device time -- the calls over a ring of input sets sized past L2 (126 MB, so no launch finds its input in cache) are captured
into one CUDA graph per side and the replays are timed with CUDA events (no Python / binding overhead on either side).  `python tools/op_sweep.py [--quick] > profiles/r02_op_sweep.txt`"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true', help='batch 32 only')
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    from oracle import build_ref_plugin
    from gan_track_b200.torch_utils import custom_ops
    from gan_track_b200.torch_utils.ops import upfirdn2d as our_upfirdn2d
    from gan_track_b200.training import networks_stylegan2 as nets
    dev = torch.device('cuda', 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_ba, ref_up = build_ref_plugin.load()
    our_ba, our_up = custom_ops.get_plugin('bias_act_plugin'), custom_ops.get_plugin('upfirdn2d_plugin')
    L2 = 126 << 20

    def ring(make, nbytes):
        n = max(2, min(8, int(math.ceil(2.5 * L2 / max(nbytes, 1)))))
        return [make() for _ in range(n)]

    def bench(fn, sets):
        """Microseconds per call on the DEVICE: the calls over the whole ring are captured into one CUDA graph and the graph is
        replayed, so that neither side pays for Python / binding overhead (the training step replays graphs too); falls back to
        eager launches if a side cannot be captured."""
        for i in range(3):
            fn(*sets[i % len(sets)])
        torch.cuda.synchronize()
        reps = max(1, args.iters // len(sets))
        try:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(*sets[0])
            torch.cuda.current_stream().wait_stream(side)
            keep = []
            with torch.cuda.graph(g):
                for x in sets:
                    keep.append(fn(*x))
            g.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                g.replay()
            e.record()
            e.synchronize()
            del keep
            return s.elapsed_time(e) / (reps * len(sets)) * 1e3
        except Exception:
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(args.iters):
                fn(*sets[i % len(sets)])
            e.record()
            e.synchronize()
            return s.elapsed_time(e) / args.iters * 1e3          # microseconds

    # layer table: resolution -> channels, fp16?   (256^2 / cbase 16384; 512^2 / cbase 32768 adds the 512^2 row)
    LAYERS = [(4, 512, False), (8, 512, False), (16, 512, False), (32, 512, True), (64, 256, True), (128, 128, True), (256, 64, True), (512, 64, True)]
    batches = [32] if args.quick else [8, 32, 64]
    f4 = our_upfirdn2d.setup_filter([1, 3, 3, 1], device=dev)
    e16, e32 = torch.empty([0], device=dev, dtype=torch.float16), torch.empty([0], device=dev, dtype=torch.float32)
    print(f'# op sweep on {torch.cuda.get_device_name(0)}; times in microseconds per call (CUDA-graph replays timed with CUDA events, inputs rotated past L2)')
    print(f'# {"op":14s} {"shape":24s} {"dtype":5s} {"layout":5s} {"reference":>10s} {"ours":>10s} {"speedup":>8s} {"ours GB/s":>10s}')

    for N in batches:
        for res, C, fp16_layer in LAYERS:
            for dtype in ([torch.float16, torch.float32] if fp16_layer else [torch.float32]):
                if res == 512 and (N > 32 or dtype == torch.float32):
                    continue                                      # 64 x 64ch x 512^2 fp32 = 4.3 GB per tensor: bounded out
                esz = 2 if dtype == torch.float16 else 4
                e = e16 if dtype == torch.float16 else e32
                shape = (N, C, res, res)
                nbytes = N * C * res * res * esz
                for cl in ([False, True] if dtype == torch.float16 else [False]):
                    mf = torch.channels_last if cl else torch.contiguous_format

                    def mk():
                        return (torch.randn(shape, device=dev).to(dtype).contiguous(memory_format=mf),)
                    sets = ring(mk, nbytes)
                    b = torch.randn([C], device=dev).to(dtype)
                    tr = bench(lambda x: ref_ba.bias_act(x, b, e, e, e, 0, 1, 3, 0.2, math.sqrt(2), 256.0), sets)
                    to = bench(lambda x: our_ba.bias_act(x, b, e, e, e, 0, 1, 3, 0.2, math.sqrt(2), 256.0), sets)
                    print(f'  {"bias_act fwd":14s} {str(shape):24s} {str(dtype)[6:]:5s} {"cl" if cl else "nchw":5s} {tr:10.1f} {to:10.1f} {tr / to:8.2f} {2 * nbytes / to / 1e3:10.0f}',
                          flush=True)
                    ys = [(x[0], ref_ba.bias_act(x[0], b, e, e, e, 0, 1, 3, 0.2, math.sqrt(2), 256.0)) for x in sets]
                    tr = bench(lambda dy, y: ref_ba.bias_act(dy, e, e, y, e, 1, 1, 3, 0.2, math.sqrt(2), 256.0), ys)
                    to = bench(lambda dy, y: our_ba.bias_act(dy, e, e, y, e, 1, 1, 3, 0.2, math.sqrt(2), 256.0), ys)
                    print(f'  {"bias_act grad1":14s} {str(shape):24s} {str(dtype)[6:]:5s} {"cl" if cl else "nchw":5s} {tr:10.1f} {to:10.1f} {tr / to:8.2f} {3 * nbytes / to / 1e3:10.0f}',
                          flush=True)
                    del ys
                    if res >= 8:
                        # the 4x4 blur after the transposed up-convolution into this resolution ([res+1]^2 -> res^2, gain 4) and the
                        # skip-branch downsample out of it
                        shp = (N, C, res + 1, res + 1)

                        def mk2():
                            return (torch.randn(shp, device=dev).to(dtype).contiguous(memory_format=mf),)
                        s2 = ring(mk2, nbytes)
                        tr = bench(lambda x: ref_up.upfirdn2d(x, f4, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0), s2)
                        to = bench(lambda x: our_up.upfirdn2d(x, f4, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0), s2)
                        print(f'  {"upfirdn blur":14s} {str(shp):24s} {str(dtype)[6:]:5s} {"cl" if cl else "nchw":5s} {tr:10.1f} {to:10.1f} {tr / to:8.2f} {2 * nbytes / to / 1e3:10.0f}',
                              flush=True)
                        del s2
                        tr = bench(lambda x: ref_up.upfirdn2d(x, f4, 1, 1, 2, 2, 1, 1, 1, 1, False, 1.0), sets)
                        to = bench(lambda x: our_up.upfirdn2d(x, f4, 1, 1, 2, 2, 1, 1, 1, 1, False, 1.0), sets)
                        print(f'  {"upfirdn down2":14s} {str(shape):24s} {str(dtype)[6:]:5s} {"cl" if cl else "nchw":5s} {tr:10.1f} {to:10.1f} {tr / to:8.2f} {1.25 * nbytes / to / 1e3:10.0f}',
                              flush=True)
                    del sets
                # one modulated 3x3 layer C -> C at this resolution (conv1 of the synthesis block) + its bias_act
                w = torch.randn([C, C, 3, 3], device=dev)
                bias = torch.randn([C], device=dev)
                fl = 2 * N * res * res * C * C * 9

                def mk3():
                    return (torch.randn(shape, device=dev).to(dtype), torch.randn([N, C], device=dev) + 1, torch.randn([N, 1, res, res], device=dev))
                s3 = ring(mk3, nbytes)

                def ref_layer(x, s, noise):
                    # S3/training/networks_stylegan2.py:52-77 (training branch) + :325-327, over the reference plugin and the library conv
                    wt, st = w, s
                    if dtype == torch.float16:
                        wt = wt * (1 / math.sqrt(C * 9) / wt.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
                        st = st / st.norm(float('inf'), dim=1, keepdim=True)
                    ww = wt.unsqueeze(0) * st.reshape(N, 1, -1, 1, 1)
                    d = (ww.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
                    y = F.conv2d(x * st.to(dtype).reshape(N, -1, 1, 1), wt.to(dtype), padding=1)
                    y = torch.addcmul(noise.to(dtype), y, d.to(dtype).reshape(N, -1, 1, 1))
                    return ref_ba.bias_act(y, bias.to(dtype), e, e, e, 0, 1, 3, 0.2, math.sqrt(2), 256.0)

                def our_layer(x, s, noise):
                    # as SynthesisLayer.forward calls it (training/networks_stylegan2.py): the layer's bias_act rides in the demodulation pass
                    return nets.modulated_conv2d(x=x, weight=w, styles=s, noise=noise, padding=1, resample_filter=f4, flip_weight=True, fused_modconv=False,
                                                 bias_act_args=dict(b=bias.to(dtype), act='lrelu', gain=None, clamp=256.0))
                if dtype == torch.float16:
                    s3 = [(x.contiguous(memory_format=torch.channels_last), s, n) for x, s, n in s3]     # the layout the package keeps fp16 blocks in
                    s3r = [(x.contiguous(), s, n) for x, s, n in s3]
                else:
                    s3r = s3
                tr = bench(ref_layer, s3r)
                to = bench(our_layer, s3)
                err = float((our_layer(*s3[0]).float() - ref_layer(*s3r[0]).float()).abs().max() / ref_layer(*s3r[0]).float().abs().max())
                print(f'  {"modconv layer":14s} {str(shape):24s} {str(dtype)[6:]:5s} {"":5s} {tr:10.1f} {to:10.1f} {tr / to:8.2f} {fl / to / 1e6:9.0f}T  rel.diff {err:.1e}',
                      flush=True)
                del s3, s3r


if __name__ == '__main__':
    main()
