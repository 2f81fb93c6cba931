// Shared problem description of the implicit-GEMM convolution kernels (conv_igemm.cu: per-tap TMA loads;
// conv_igemm_halo.cu: halo-staged persistent kernel).
#pragma once
#include "gt_common.cuh"
#include "gt_sm100.cuh"
#include "hot_act.cuh"

constexpr int CONV_MAX_TAPS = 9;
constexpr int CONV_MAX_PHASES = 4;

// One output "phase": a stride-1 (or input-strided) correlation with its own tap list.  A plain convolution has one
// phase; a stride-2 transposed convolution has four (output parity), each writing every other output pixel.
struct ConvPhase {
    int ntaps;
    int OHp, OWp;        // extent of this phase's output grid
    int off_y, off_x;    // output pixel = (a * out_stride + off_y, b * out_stride + off_x)
    long long y_off;     // element offset of this phase's output slab (K-split phases of the fp32-output mode write partial sums side by side)
    int8_t tdy[CONV_MAX_TAPS], tdx[CONV_MAX_TAPS], tw[CONV_MAX_TAPS];   // input offset of the tap, weight slab of the tap
};

struct ConvParams {
    ConvPhase ph[CONV_MAX_PHASES];
    int nphases;
    int N, Cin, Cout;
    int in_stride, out_stride;
    int bw_log2, bh_log2;            // (per-tap kernel) tile box bw x bh x bn pixels, product 128
    int tiles_w, tiles_h, tiles_n;   // over the largest phase
    int n_tiles;                     // Cout / BN
    void* y;                         // fp16 (f32out == 0) or fp32 (f32out == 1) NHWC output
    long long ys_n, ys_h, ys_w;      // element strides
    int f32out;                      // 1: raw fp32 accumulators are stored (fp16x3 route of the fp32 layers, csrc/conv_f16x3.cu); operands stay fp16
    // optional fused epilogue (fp16 output only): y = clamp(act(round_fp16(acc) + bias[co]) * gain), the bias_act that follows the
    // convolution in Conv2dLayer.forward (S3/training/networks_stylegan2.py:176-177); ep_act 0 = plain store
    const __half* ep_bias;
    const __half* ep_add;            // optional residual with the layout of y: y = fp16(epilogue) + ep_add (DiscriminatorBlock's y.add_(x), :636)
    int ep_act;
    float ep_alpha, ep_gain, ep_clamp;
    int dbg_no_store;                // experiments only: run the whole pipeline but skip the global stores
    // halo kernel only
    int dy_min[CONV_MAX_PHASES], dx_min[CONV_MAX_PHASES];   // most negative tap offset of the phase = halo origin
    int halo_w, halo_h;                                      // TMA box extent in pixels
};

// Kernel flavours (compile-time, so that the plain fp16 path carries none of the others' code or registers):
enum { CONV_F16 = 0, CONV_F32OUT = 1, CONV_F16_EP = 2 };
__host__ __device__ inline int conv_mode(const ConvParams& p) { return p.f32out ? CONV_F32OUT : (p.ep_act != 0 ? CONV_F16_EP : CONV_F16); }

// 32 consecutive accumulator columns (output channels co0 .. co0+31) of one output pixel -> global memory
// (fp16: 64 bytes, fp32: 128 bytes); CONV_F16_EP runs the bias + activation epilogue on the way
template <int MODE>
__device__ __forceinline__ void conv_store32(const ConvParams& p, long long elem_off, int co0, const uint32_t (&r)[32]) {
    if (MODE == CONV_F32OUT) {
        float* yp = (float*)p.y + elem_off;
#pragma unroll
        for (int v = 0; v < 8; v++) *reinterpret_cast<uint4*>(yp + v * 4) = make_uint4(r[v * 4], r[v * 4 + 1], r[v * 4 + 2], r[v * 4 + 3]);
        return;
    }
    __half* yp = (__half*)p.y + elem_off;
#pragma unroll
    for (int v = 0; v < 4; v++) {
        __half2 h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) h[k] = __floats2half2_rn(__uint_as_float(r[v * 8 + 2 * k]), __uint_as_float(r[v * 8 + 2 * k + 1]));
        if (MODE == CONV_F16_EP) {
            // the reference materialises the convolution output in fp16 before bias_act reads it back: round first
            const hot::Params hp = hot::make_params(p.ep_alpha, p.ep_gain, p.ep_clamp);
            const bool clamp_on = p.ep_clamp >= 0.f;
            Vec16<__half> bv;
            if (p.ep_bias) bv = ld16(p.ep_bias + co0 + v * 8);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float2 u = __half22float2(h[k]);
                if (p.ep_bias) u = __fadd2_rn(u, __half22float2(reinterpret_cast<const __half2*>(bv.v)[k]));
                float2 o;
                if (p.ep_act == hot::LRELU) o = clamp_on ? hot::fwd<hot::LRELU, true>(u, hp) : hot::fwd<hot::LRELU, false>(u, hp);
                else o = clamp_on ? hot::fwd<hot::LINEAR, true>(u, hp) : hot::fwd<hot::LINEAR, false>(u, hp);
                h[k] = __float22half2_rn(o);
            }
            if (p.ep_add) {                                   // fp16 + fp16 -> fp16, round to nearest: what Tensor.add_ computes
                const Vec16<__half> av = ld16(p.ep_add + elem_off + v * 8);
#pragma unroll
                for (int k = 0; k < 4; k++) h[k] = __hadd2(h[k], reinterpret_cast<const __half2*>(av.v)[k]);
            }
        }
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&h[0]);
        o.y = *reinterpret_cast<uint32_t*>(&h[1]);
        o.z = *reinterpret_cast<uint32_t*>(&h[2]);
        o.w = *reinterpret_cast<uint32_t*>(&h[3]);
        *reinterpret_cast<uint4*>(yp + v * 8) = o;
    }
}

// conv_igemm_halo.cu
int gt_launch_conv_halo(const void* x, long long xs_n, long long xs_h, long long xs_w, int H, int W, const void* wpacked, int ntaps_total, ConvParams& p,
                        cudaStream_t stream);
bool gt_conv_halo_applicable(const ConvParams& p, int maxOH, int maxOW);

// conv_rows.cu: row-streaming kernel for the 64 -> 64 channel 3x3 stride-1 layers
bool gt_conv_rows_applicable(const ConvParams& p, int H, int W);
int gt_launch_conv_rows(const void* x, long long xs_n, long long xs_h, long long xs_w, int H, int W, const void* wpacked, int ntaps_total, const ConvParams& p,
                        cudaStream_t stream);
