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

// ---------------------------------------------------------------------------------------------------------------------
// Two-pass form (what the entry points launch).  The single-gather kernels above evaluate 49 taps per output pixel and
// are issue-bound (ncu, profiles/r01b: 310 M warp instructions, 870 us for [32,1,256,256] -> 524x524); the separable
// structure is cheaper when the 2x-upsampled image is materialised once -- still with device-resident margins:
//     pass A  U = upsample2x(reflect_pad(x))        CTA tile 32x32 of U from a 22x22 source tile, row pass then column pass
//     pass B  y = bilinear(U, affine grid), zeros outside U                                        4 taps per output
// U lives in a caller-provided workspace sized for the largest margins (gt_aug_warp_workspace) and is laid out compactly
// with the ACTUAL extents (pitch wu = 2 (W + mx0 + mx1)) computed in-kernel from the margins, so only the used part is
// ever touched.  Backward is the adjoint of each pass in reverse order (B^T scatters with atomicAdd into a zeroed gU like
// aten's grid_sampler_2d_backward; A^T = 12-tap stride-2 correlation per axis, then the reflect fold with atomicAdd).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int UT = 32;                       // U tile edge
constexpr int SRC_MAX = UT / 2 + 6;          // source rows/cols per U tile for <= 12 taps
constexpr int GT = 16;                       // padded-image tile edge of the adjoint upsampling pass
constexpr int GU_MAX = 2 * (GT - 1) + 12;    // gU rows/cols per tile for <= 12 taps

struct Extents {
    int mx0, my0, mx1, my1, npx, npy, wu, hu;
};
__device__ __forceinline__ Extents extents(const WarpParams& p) {
    Extents e;
    e.mx0 = p.margins[0];
    e.my0 = p.margins[1];
    e.mx1 = p.margins[2];
    e.my1 = p.margins[3];
    e.npx = p.W + e.mx0 + e.mx1;
    e.npy = p.H + e.my0 + e.my1;
    e.wu = 2 * e.npx;
    e.hu = 2 * e.npy;
    return e;
}
__device__ __forceinline__ int reflect_idx(int s, int n) {
    if (s < 0) s = -s;
    if (s >= n) s = 2 * (n - 1) - s;
    return s;
}

// pass A.  U[Y] = 2 * sum_k f[q - 2k] * P[i0 + k],  i0 = ceil((Y - pad0) / 2),  q = base + Y - 2 i0   (see axis_weights)
__global__ void __launch_bounds__(256) aug_up_kernel(const WarpParams p, float* __restrict__ U) {
    __shared__ float sP[SRC_MAX][SRC_MAX + 1];
    __shared__ float sT[SRC_MAX][UT + 1];
    __shared__ float sf[WARP_MAX_TAPS];
    if (threadIdx.x < WARP_MAX_TAPS) sf[threadIdx.x] = threadIdx.x < p.ntaps ? p.taps[threadIdx.x] : 0.f;
    const Extents e = extents(p);
    const int T = p.ntaps, HT = T / 2, pad0 = (T + 1) / 2, base = T - 1 - pad0;
    const int src = UT / 2 + HT;
    const int tiles_x = (e.wu + UT - 1) / UT, tiles_y = (e.hu + UT - 1) / UT;
    const long long total = (long long)p.B * p.C * tiles_y * tiles_x;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long r0 = tile / tiles_x;
        const int ty = (int)(r0 % tiles_y);
        const long long plane = r0 / tiles_y;
        const int X0 = tx * UT, Y0 = ty * UT;
        const int jb = (X0 - pad0 + 1) >> 1, ib = (Y0 - pad0 + 1) >> 1;
        const float* xp = p.x + plane * p.H * p.W;
        __syncthreads();
        for (int t = threadIdx.x; t < src * src; t += 256) {
            const int r = t / src, c = t - r * src;
            const int i = ib + r, j = jb + c;
            float v = 0.f;
            if (i >= 0 && i < e.npy && j >= 0 && j < e.npx) v = __ldg(xp + (long long)reflect_idx(i - e.my0, p.H) * p.W + reflect_idx(j - e.mx0, p.W));
            sP[r][c] = v;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < src * UT; t += 256) {
            const int r = t / UT, xx = t - r * UT;
            const int X = X0 + xx;
            const int i0 = (X - pad0 + 1) >> 1, q = base + X - 2 * i0;
            float acc = 0.f;
            for (int k = 0; k < HT; k++) acc += sf[q - 2 * k] * sP[r][i0 - jb + k];
            sT[r][xx] = 2.f * acc;
        }
        __syncthreads();
        float* up = U + plane * ((long long)e.hu * e.wu);
        for (int t = threadIdx.x; t < UT * UT; t += 256) {
            const int yy = t / UT, xx = t - yy * UT;
            const int Y = Y0 + yy, X = X0 + xx;
            const int i0 = (Y - pad0 + 1) >> 1, q = base + Y - 2 * i0;
            float acc = 0.f;
            for (int k = 0; k < HT; k++) acc += sf[q - 2 * k] * sT[i0 - ib + k][xx];
            if (Y < e.hu && X < e.wu) up[(long long)Y * e.wu + X] = 2.f * acc;
        }
    }
}

// bilinear footprint of one output pixel in U
struct Bilin {
    int X0, Y0;
    float wx0, wx1, wy0, wy1;     // already zeroed for taps outside U
};
__device__ __forceinline__ Bilin bilin(const WarpParams& p, const Extents& e, int b, int oy, int ox) {
    const float* th = p.theta + b * 6;
    // coordinates in double, as in footprint() above
    const double xn = (2.0 * ox + 1.0) / p.OW - 1.0;
    const double yn = (2.0 * oy + 1.0) / p.OH - 1.0;
    const double gx = (double)th[0] * xn + (double)th[1] * yn + (double)th[2];
    const double gy = (double)th[3] * xn + (double)th[4] * yn + (double)th[5];
    const double ix = ((gx + 1.0) * (double)e.wu - 1.0) * 0.5;
    const double iy = ((gy + 1.0) * (double)e.hu - 1.0) * 0.5;
    const double fx = floor(ix), fy = floor(iy);
    Bilin r;
    // clamp far-out coordinates before the int conversion; anything beyond one pixel outside contributes nothing
    r.X0 = (int)fmax(fmin(fx, 1.0e9), -1.0e9);
    r.Y0 = (int)fmax(fmin(fy, 1.0e9), -1.0e9);
    const float ax = (float)(ix - fx), ay = (float)(iy - fy);
    r.wx0 = (r.X0 >= 0 && r.X0 < e.wu) ? 1.f - ax : 0.f;
    r.wx1 = (r.X0 + 1 >= 0 && r.X0 + 1 < e.wu) ? ax : 0.f;
    r.wy0 = (r.Y0 >= 0 && r.Y0 < e.hu) ? 1.f - ay : 0.f;
    r.wy1 = (r.Y0 + 1 >= 0 && r.Y0 + 1 < e.hu) ? ay : 0.f;
    return r;
}

// pass B
__global__ void __launch_bounds__(256) aug_sample_kernel(const WarpParams p, const float* __restrict__ U) {
    const Extents e = extents(p);
    const long long total = (long long)p.B * p.OH * p.OW;
    const long long plane_sz = (long long)e.hu * e.wu;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % p.OW);
        const long long t = idx / p.OW;
        const int oy = (int)(t % p.OH);
        const int b = (int)(t / p.OH);
        const Bilin w = bilin(p, e, b, oy, ox);
        for (int c = 0; c < p.C; c++) {
            const float* up = U + ((long long)b * p.C + c) * plane_sz;
            float acc = 0.f;
            if (w.wy0 != 0.f) {
                const float* row = up + (long long)w.Y0 * e.wu;
                float ra = 0.f;
                if (w.wx0 != 0.f) ra += w.wx0 * __ldg(row + w.X0);
                if (w.wx1 != 0.f) ra += w.wx1 * __ldg(row + w.X0 + 1);
                acc += w.wy0 * ra;
            }
            if (w.wy1 != 0.f) {
                const float* row = up + (long long)(w.Y0 + 1) * e.wu;
                float ra = 0.f;
                if (w.wx0 != 0.f) ra += w.wx0 * __ldg(row + w.X0);
                if (w.wx1 != 0.f) ra += w.wx1 * __ldg(row + w.X0 + 1);
                acc += w.wy1 * ra;
            }
            p.y[(((long long)b * p.C + c) * p.OH + oy) * p.OW + ox] = acc;
        }
    }
}

// zero the used part of gU (its size is only known on the device)
__global__ void __launch_bounds__(256) aug_zero_kernel(const WarpParams p, float* __restrict__ gU) {
    const Extents e = extents(p);
    const long long total = (long long)p.B * p.C * e.hu * e.wu;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += (long long)gridDim.x * blockDim.x * 4) {
        if (i + 3 < total) *reinterpret_cast<float4*>(gU + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        else for (long long k = i; k < total; k++) gU[k] = 0.f;
    }
}

// adjoint of pass B: p.y holds the incoming gradient
__global__ void __launch_bounds__(256) aug_sample_adj_kernel(const WarpParams p, float* __restrict__ gU) {
    const Extents e = extents(p);
    const long long total = (long long)p.B * p.OH * p.OW;
    const long long plane_sz = (long long)e.hu * e.wu;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % p.OW);
        const long long t = idx / p.OW;
        const int oy = (int)(t % p.OH);
        const int b = (int)(t / p.OH);
        const Bilin w = bilin(p, e, b, oy, ox);
        for (int c = 0; c < p.C; c++) {
            const float g = p.y[(((long long)b * p.C + c) * p.OH + oy) * p.OW + ox];
            if (g == 0.f) continue;
            float* up = gU + ((long long)b * p.C + c) * plane_sz;
            if (w.wy0 != 0.f) {
                float* row = up + (long long)w.Y0 * e.wu;
                if (w.wx0 != 0.f) atomicAdd(row + w.X0, g * w.wy0 * w.wx0);
                if (w.wx1 != 0.f) atomicAdd(row + w.X0 + 1, g * w.wy0 * w.wx1);
            }
            if (w.wy1 != 0.f) {
                float* row = up + (long long)(w.Y0 + 1) * e.wu;
                if (w.wx0 != 0.f) atomicAdd(row + w.X0, g * w.wy1 * w.wx0);
                if (w.wx1 != 0.f) atomicAdd(row + w.X0 + 1, g * w.wy1 * w.wx1);
            }
        }
    }
}

// adjoint of pass A: gP[i][j] = sum_{Y,X} 2 f[base + Y - 2i] * 2 f[base + X - 2j] * gU[Y][X]; then gx[refl(i), refl(j)] += gP[i][j]
__global__ void __launch_bounds__(256) aug_up_adj_kernel(const WarpParams p, const float* __restrict__ gU, float* __restrict__ gx) {
    __shared__ float sG[GU_MAX][GU_MAX + 1];
    __shared__ float sT[GU_MAX][GT + 1];
    __shared__ float sf[WARP_MAX_TAPS];
    if (threadIdx.x < WARP_MAX_TAPS) sf[threadIdx.x] = threadIdx.x < p.ntaps ? p.taps[threadIdx.x] : 0.f;
    const Extents e = extents(p);
    const int T = p.ntaps, pad0 = (T + 1) / 2, base = T - 1 - pad0;
    const int ext = 2 * (GT - 1) + T;
    const int tiles_x = (e.npx + GT - 1) / GT, tiles_y = (e.npy + GT - 1) / GT;
    const long long total = (long long)p.B * p.C * tiles_y * tiles_x;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long r0 = tile / tiles_x;
        const int ty = (int)(r0 % tiles_y);
        const long long plane = r0 / tiles_y;
        const int jb = tx * GT, ib = ty * GT;
        const int Xb = 2 * jb - base, Yb = 2 * ib - base;          // first gU row / column any tap of this tile reads
        const float* up = gU + plane * ((long long)e.hu * e.wu);
        __syncthreads();
        for (int t = threadIdx.x; t < ext * ext; t += 256) {
            const int r = t / ext, c = t - r * ext;
            const int Y = Yb + r, X = Xb + c;
            sG[r][c] = (Y >= 0 && Y < e.hu && X >= 0 && X < e.wu) ? __ldg(up + (long long)Y * e.wu + X) : 0.f;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < ext * GT; t += 256) {        // row pass: T[r][j] = sum_t 2 f[t] * G[r][2 (j - jb) + t]
            const int r = t / GT, jj = t - r * GT;
            float acc = 0.f;
            for (int k = 0; k < T; k++) acc += sf[k] * sG[r][2 * jj + k];
            sT[r][jj] = 2.f * acc;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < GT * GT; t += 256) {
            const int ii = t / GT, jj = t - ii * GT;
            const int i = ib + ii, j = jb + jj;
            if (i < e.npy && j < e.npx) {
                float acc = 0.f;
                for (int k = 0; k < T; k++) acc += sf[k] * sT[2 * ii + k][jj];
                atomicAdd(gx + plane * p.H * p.W + (long long)reflect_idx(i - e.my0, p.H) * p.W + reflect_idx(j - e.mx0, p.W), 2.f * acc);
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

// floats of workspace for the 2x-upsampled image at the largest margins the caller may pass (each in [0, W-1] / [0, H-1])
extern "C" long long gt_aug_warp_workspace(int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return (long long)B * C * (2ll * (3 * H - 2)) * (2ll * (3 * W - 2));
}

extern "C" int gt_aug_warp_fwd(const float* x, const float* theta, const int* margins, const float* taps_host, int ntaps, float* y, int B, int C, int H,
                               int W, int OH, int OW, float* workspace, long long workspace_floats, void* stream) {
    GT_REQUIRE(x != nullptr, "gt_aug_warp_fwd: null pointer");
    WarpParams p;
    int rc = fill_params(p, x, theta, margins, taps_host, ntaps, y, B, C, H, W, OH, OW);
    if (rc != GT_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (workspace == nullptr) {      // single-gather form (no workspace)
        aug_warp_fwd_kernel<<<grid_for((long long)B * OH * OW), 256, 0, st>>>(p);
        GT_CUDA_LAUNCH_CHECK("gt_aug_warp_fwd");
        return GT_OK;
    }
    GT_REQUIRE(workspace_floats >= gt_aug_warp_workspace(B, C, H, W), "gt_aug_warp_fwd: workspace too small");
    aug_up_kernel<<<gt_num_sms() * 8, 256, 0, st>>>(p, workspace);
    GT_CUDA_LAUNCH_CHECK("gt_aug_warp_fwd(upsample)");
    aug_sample_kernel<<<grid_for((long long)B * OH * OW), 256, 0, st>>>(p, workspace);
    GT_CUDA_LAUNCH_CHECK("gt_aug_warp_fwd(sample)");
    return GT_OK;
}

extern "C" int gt_aug_warp_bwd(const float* gy, const float* theta, const int* margins, const float* taps_host, int ntaps, float* gx, int B, int C, int H,
                               int W, int OH, int OW, float* workspace, long long workspace_floats, void* stream) {
    GT_REQUIRE(gy != nullptr && gx != nullptr, "gt_aug_warp_bwd: null pointer");
    WarpParams p;
    int rc = fill_params(p, nullptr, theta, margins, taps_host, ntaps, const_cast<float*>(gy), B, C, H, W, OH, OW);
    if (rc != GT_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)B * C * H * W, st);
    if (e != cudaSuccess) {
        gt_set_error("gt_aug_warp_bwd: memset failed: %s", cudaGetErrorString(e));
        return GT_ERR_CUDA;
    }
    if (workspace == nullptr) {
        aug_warp_bwd_kernel<<<grid_for((long long)B * OH * OW), 256, 0, st>>>(p, gx);
        GT_CUDA_LAUNCH_CHECK("gt_aug_warp_bwd");
        return GT_OK;
    }
    GT_REQUIRE(workspace_floats >= gt_aug_warp_workspace(B, C, H, W), "gt_aug_warp_bwd: workspace too small");
    aug_zero_kernel<<<gt_num_sms() * 8, 256, 0, st>>>(p, workspace);
    GT_CUDA_LAUNCH_CHECK("gt_aug_warp_bwd(zero)");
    aug_sample_adj_kernel<<<grid_for((long long)B * OH * OW), 256, 0, st>>>(p, workspace);
    GT_CUDA_LAUNCH_CHECK("gt_aug_warp_bwd(sample^T)");
    aug_up_adj_kernel<<<gt_num_sms() * 8, 256, 0, st>>>(p, workspace, gx);
    GT_CUDA_LAUNCH_CHECK("gt_aug_warp_bwd(upsample^T)");
    return GT_OK;
}
