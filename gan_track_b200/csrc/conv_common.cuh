// Shared problem description of the implicit-GEMM convolution kernels (conv_igemm.cu: per-tap TMA loads;
// conv_igemm_halo.cu: halo-staged persistent kernel).
#pragma once
#include "gt_common.cuh"
#include "gt_sm100.cuh"

constexpr int CONV_MAX_TAPS = 9;
constexpr int CONV_MAX_PHASES = 4;

// One output "phase": a stride-1 (or input-strided) correlation with its own tap list.  A plain convolution has one
// phase; a stride-2 transposed convolution has four (output parity), each writing every other output pixel.
struct ConvPhase {
    int ntaps;
    int OHp, OWp;        // extent of this phase's output grid
    int off_y, off_x;    // output pixel = (a * out_stride + off_y, b * out_stride + off_x)
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
    void* y;                         // fp16 (tf32 == 0) or fp32 (tf32 == 1) NHWC output
    long long ys_n, ys_h, ys_w;      // element strides
    int tf32;                        // 1: operands are fp32 bit patterns consumed as TF32 (32 channels per 128-byte chunk), fp32 output
    // halo kernel only
    int dy_min[CONV_MAX_PHASES], dx_min[CONV_MAX_PHASES];   // most negative tap offset of the phase = halo origin
    int halo_w, halo_h;                                      // TMA box extent in pixels
};

// 32 consecutive accumulator columns of one output pixel -> global memory (fp16: 64 bytes, fp32: 128 bytes)
__device__ __forceinline__ void conv_store32(void* y, long long elem_off, int tf32, const uint32_t (&r)[32]) {
    if (tf32) {
        float* yp = (float*)y + elem_off;
#pragma unroll
        for (int v = 0; v < 8; v++) *reinterpret_cast<uint4*>(yp + v * 4) = make_uint4(r[v * 4], r[v * 4 + 1], r[v * 4 + 2], r[v * 4 + 3]);
    } else {
        __half* yp = (__half*)y + elem_off;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            uint4 o;
            __half2 h0 = __floats2half2_rn(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
            __half2 h1 = __floats2half2_rn(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
            __half2 h2 = __floats2half2_rn(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
            __half2 h3 = __floats2half2_rn(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
            o.x = *reinterpret_cast<uint32_t*>(&h0);
            o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2);
            o.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(yp + v * 8) = o;
        }
    }
}

// conv_igemm_halo.cu
int gt_launch_conv_halo(const void* x, long long xs_n, long long xs_h, long long xs_w, int H, int W, const void* wpacked, int ntaps_total, ConvParams& p,
                        cudaStream_t stream);
bool gt_conv_halo_applicable(const ConvParams& p, int maxOH, int maxOW);
