// augment_warp.cu -- the geometric core of the ADA pipe as ONE gather kernel (and its adjoint), with no device->host
// synchronisation.
//
// Reference (S3/training/augment_mi.py:286-318, S3 = /root/reference/src/models/stylegan3):
//     margins = ceil(max over the batch of the warped corners ...)          -> four Python ints (a .cpu() sync, :289-299)
//     images  = F.pad(images, margins, mode='reflect')                                                        (:303)
//     images  = upfirdn2d.upsample2d(images, Hz_geom, up=2)                 12-tap sym6, two separable passes (:310)
//     images  = grid_sample(images, affine_grid(theta, [B,C,2(H+6),2(W+6)]), bilinear, zeros, align_corners=False) (:315-318)
// Every output pixel of the sampled grid is a linear function of at most 7x7 pixels of the ORIGINAL image: the bilinear
// footprint (2x2) of the 2x-upsampled image pulls in two polyphase components of the 12-tap filter per axis, i.e. 7
// source rows x 7 source columns with separable weights, and the reflect padding is an index map.  So the padded and
// the upsampled images are never materialised; the margins stay on the device (an int[4] the kernel reads), which is
// what makes the training step capturable in a CUDA graph.  Work: 49 FMAs per output pixel; the 1-channel 256x256
// inputs live in L2.
//
// Exact semantics kept: zeros outside the padded image (upfirdn2d zero padding), zeros outside the upsampled image
// (grid_sample padding_mode='zeros'), reflect without edge repeat, margins clamped by the caller to [0, W-1].
// The op is linear in the image, so backward is the adjoint (scatter with atomicAdd, as aten's own
// grid_sampler_2d_backward does) and the double backward (R1 penalty) is the forward kernel again.
#include "gt_common.cuh"

namespace {

constexpr int WARP_MAX_TAPS = 16;

struct WarpParams {
    const float* x;        // [B,C,H,W] contiguous
    const float* theta;    // [B,2,3]
    const int* margins;    // device int[4]: mx0, my0, mx1, my1
    float* y;              // [B,C,OH,OW]
    int B, C, H, W, OH, OW, ntaps;
    float taps[WARP_MAX_TAPS];   // low-pass taps as upfirdn2d.setup_filter stores them (normalised, not flipped)
};

struct Footprint {
    float wy[7], wx[7];
    int ry[7], rx[7];      // source row / column (reflect already applied) or -1
};

// Weights of the (up to) 7 source samples along one axis for the bilinear sample at upsampled coordinate `pos`.
// U[Y] = 2 * sum_i f[5 + Y - 2i] * P[i]  (upfirdn2d up=2, pad (6,5), gain 2 per axis, true convolution, 12 taps; for a
// general even tap count T: pad0 = (T+1)/2, U[Y] = 2 * sum_i f[T-1-pad0 + Y - 2i] * P[i]).
__device__ __forceinline__ void axis_weights(double pos, int n_src, int m0, int m1, const float* f, int T, float* w, int* r) {
    const int np_ = n_src + m0 + m1;                 // padded extent
    const int nu = 2 * np_;                          // upsampled extent
    const double fl = floor(pos);
    const float frac = (float)(pos - fl);
    const int Y0 = (int)fl;
    const float a0 = (Y0 >= 0 && Y0 < nu) ? (1.f - frac) : 0.f;
    const float a1 = (Y0 + 1 >= 0 && Y0 + 1 < nu) ? frac : 0.f;
    const int pad0 = (T + 1) / 2;
    const int base = T - 1 - pad0;                   // 5 for T = 12
    const int i0 = (Y0 - pad0 + 1) >> 1;             // ceil((Y0 - pad0) / 2), arithmetic shift = floor
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const int i = i0 + k;
        const int t0 = base + Y0 - 2 * i, t1 = t0 + 1;
        float wgt = 0.f;
        if (t0 >= 0 && t0 < T) wgt += a0 * f[t0];
        if (t1 >= 0 && t1 < T) wgt += a1 * f[t1];
        int src = -1;
        if (i >= 0 && i < np_ && wgt != 0.f) {
            int s = i - m0;
            if (s < 0) s = -s;
            if (s >= n_src) s = 2 * (n_src - 1) - s;
            src = s;
        }
        w[k] = (src >= 0) ? 2.f * wgt : 0.f;
        r[k] = src;
    }
}

__device__ __forceinline__ void footprint(const WarpParams& p, int b, int oy, int ox, Footprint& fp) {
    const float* th = p.theta + b * 6;
    const int mx0 = p.margins[0], my0 = p.margins[1], mx1 = p.margins[2], my1 = p.margins[3];
    // Coordinates in double: the reference evaluates them in fp32 in a different association order (linspace, matmul,
    // unnormalise), which alone is worth ~3e-5 px at 512-px extents; exact arithmetic here halves the mismatch.
    const double xn = (2.0 * ox + 1.0) / p.OW - 1.0;
    const double yn = (2.0 * oy + 1.0) / p.OH - 1.0;
    const double gx = (double)th[0] * xn + (double)th[1] * yn + (double)th[2];
    const double gy = (double)th[3] * xn + (double)th[4] * yn + (double)th[5];
    const double wu = 2.0 * (p.W + mx0 + mx1), hu = 2.0 * (p.H + my0 + my1);
    const double ix = ((gx + 1.0) * wu - 1.0) * 0.5;
    const double iy = ((gy + 1.0) * hu - 1.0) * 0.5;
    axis_weights(ix, p.W, mx0, mx1, p.taps, p.ntaps, fp.wx, fp.rx);
    axis_weights(iy, p.H, my0, my1, p.taps, p.ntaps, fp.wy, fp.ry);
}

__global__ void __launch_bounds__(256) aug_warp_fwd_kernel(const WarpParams p) {
    const long long total = (long long)p.B * p.OH * p.OW;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % p.OW);
        const long long t = idx / p.OW;
        const int oy = (int)(t % p.OH);
        const int b = (int)(t / p.OH);
        Footprint fp;
        footprint(p, b, oy, ox, fp);
        for (int c = 0; c < p.C; c++) {
            const float* xp = p.x + ((long long)b * p.C + c) * p.H * p.W;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 7; i++) {
                if (fp.ry[i] < 0) continue;
                const float* row = xp + (long long)fp.ry[i] * p.W;
                float racc = 0.f;
#pragma unroll
                for (int j = 0; j < 7; j++)
                    if (fp.rx[j] >= 0) racc += fp.wx[j] * __ldg(row + fp.rx[j]);
                acc += fp.wy[i] * racc;
            }
            p.y[(((long long)b * p.C + c) * p.OH + oy) * p.OW + ox] = acc;
        }
    }
}

// adjoint: p.y is the incoming gradient [B,C,OH,OW] (read), p.x is reinterpreted as the OUTPUT gradient (atomicAdd target)
__global__ void __launch_bounds__(256) aug_warp_bwd_kernel(const WarpParams p, float* __restrict__ gx) {
    const long long total = (long long)p.B * p.OH * p.OW;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % p.OW);
        const long long t = idx / p.OW;
        const int oy = (int)(t % p.OH);
        const int b = (int)(t / p.OH);
        Footprint fp;
        footprint(p, b, oy, ox, fp);
        for (int c = 0; c < p.C; c++) {
            const float g = p.y[(((long long)b * p.C + c) * p.OH + oy) * p.OW + ox];
            if (g == 0.f) continue;
            float* gp = gx + ((long long)b * p.C + c) * p.H * p.W;
#pragma unroll
            for (int i = 0; i < 7; i++) {
                if (fp.ry[i] < 0) continue;
                const float gi = g * fp.wy[i];
                float* row = gp + (long long)fp.ry[i] * p.W;
#pragma unroll
                for (int j = 0; j < 7; j++)
                    if (fp.rx[j] >= 0) atomicAdd(row + fp.rx[j], gi * fp.wx[j]);
            }
        }
    }
}

int fill_params(WarpParams& p, const float* x, const float* theta, const int* margins, const float* taps_host, int ntaps, float* y, int B, int C, int H,
                int W, int OH, int OW) {
    GT_REQUIRE(theta && margins && taps_host && y, "gt_aug_warp: null pointer");
    GT_REQUIRE(ntaps >= 2 && ntaps <= WARP_MAX_TAPS && ntaps % 2 == 0 && ntaps <= 12, "gt_aug_warp: %d taps not supported (even, <= 12)", ntaps);
    GT_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1 && OH > 0 && OW > 0, "gt_aug_warp: bad shape");
    memset(&p, 0, sizeof(p));
    p.x = x;
    p.theta = theta;
    p.margins = margins;
    p.y = y;
    p.B = B;
    p.C = C;
    p.H = H;
    p.W = W;
    p.OH = OH;
    p.OW = OW;
    p.ntaps = ntaps;
    for (int i = 0; i < ntaps; i++) p.taps[i] = taps_host[i];
    return GT_OK;
}

int grid_for(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)gt_num_sms() * 16;
    return (int)(g < cap ? g : cap);
}

}  // namespace

extern "C" int gt_aug_warp_fwd(const float* x, const float* theta, const int* margins, const float* taps_host, int ntaps, float* y, int B, int C, int H,
                               int W, int OH, int OW, void* stream) {
    GT_REQUIRE(x != nullptr, "gt_aug_warp_fwd: null pointer");
    WarpParams p;
    int rc = fill_params(p, x, theta, margins, taps_host, ntaps, y, B, C, H, W, OH, OW);
    if (rc != GT_OK) return rc;
    aug_warp_fwd_kernel<<<grid_for((long long)B * OH * OW), 256, 0, (cudaStream_t)stream>>>(p);
    GT_CUDA_LAUNCH_CHECK("gt_aug_warp_fwd");
    return GT_OK;
}

extern "C" int gt_aug_warp_bwd(const float* gy, const float* theta, const int* margins, const float* taps_host, int ntaps, float* gx, int B, int C, int H,
                               int W, int OH, int OW, void* stream) {
    GT_REQUIRE(gy != nullptr && gx != nullptr, "gt_aug_warp_bwd: null pointer");
    WarpParams p;
    int rc = fill_params(p, nullptr, theta, margins, taps_host, ntaps, const_cast<float*>(gy), B, C, H, W, OH, OW);
    if (rc != GT_OK) return rc;
    cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)B * C * H * W, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        gt_set_error("gt_aug_warp_bwd: memset failed: %s", cudaGetErrorString(e));
        return GT_ERR_CUDA;
    }
    aug_warp_bwd_kernel<<<grid_for((long long)B * OH * OW), 256, 0, (cudaStream_t)stream>>>(p, gx);
    GT_CUDA_LAUNCH_CHECK("gt_aug_warp_bwd");
    return GT_OK;
}
