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
    __half* y;
    long long ys_n, ys_h, ys_w;
    // halo kernel only
    int dy_min[CONV_MAX_PHASES], dx_min[CONV_MAX_PHASES];   // most negative tap offset of the phase = halo origin
    int halo_w, halo_h;                                      // TMA box extent in pixels
};

// conv_igemm_halo.cu
int gt_launch_conv_halo(const void* x, long long xs_n, long long xs_h, long long xs_w, int H, int W, const void* wpacked, int ntaps_total, ConvParams& p,
                        cudaStream_t stream);
bool gt_conv_halo_applicable(const ConvParams& p, int maxOH, int maxOW);
