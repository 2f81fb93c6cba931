// upfirdn2d for sm_100a: zero-insert upsample -> pad/crop -> FIR -> decimate, per channel.
//
// Replaces the reference plugin OPS/upfirdn2d.{cpp,cu} (launcher upfirdn2d.cpp:16-98; kernels upfirdn2d.cu:29-200;
// OPS = /root/reference/src/models/stylegan3/torch_utils/ops).  Contract kept: arbitrary strides on x and y, fp32
// taps, fp16/fp32/fp64 I/O with fp32 (fp64) accumulation, out = (in*up + pad0 + pad1 - f + down) / down.
//
//   y[n,c,oy,ox] = gain * sum_{ky,kx} g[ky,kx] * P[oy*downy + ky, ox*downx + kx]
//   P = x with (up-1) zeros after every sample, padded by (padx0, pady0); g = f if flip else f reversed.
//
// Only taps that land on real samples are visited: ky starts at the first tap whose padded row is a multiple of upy
// and advances by upy, so the inner loops do fh/upy x fw/upx multiply-adds.  Two kernels:
//   * upfirdn2d_vec_kernel  - channels-last tensors: one thread = one output pixel x one 16-byte channel vector.
//   * upfirdn2d_row_kernel  - W-contiguous tensors (NCHW): one thread = one output pixel, consecutive threads walk ox.
// Both take the up/down/filter-size as template parameters for the shapes on the StyleGAN2 path (4x4 taps with
// up/down in {1,2}; 12-tap separable passes of the ADA pipe) and fall back to run-time loops otherwise.
#include "gt_common.cuh"

int gt_upfirdn2d_try_tma(const void* x, const float* f, long long fs_h, long long fs_w, int flip, float gain, void* y, int dtype, int N, int C, int H, int W,
                         long long xs_n, long long xs_h, long long xs_w, int OH, int OW, long long ys_n, long long ys_h, long long ys_w, int padx0, int pady0,
                         cudaStream_t st);
int gt_upfirdn2d_try_tma_strided(const void* x, const float* f, long long fs_h, long long fs_w, int flip, float gain, void* y, int dtype, int N, int C, int H, int W,
                                 long long xs_n, long long xs_h, long long xs_w, int OH, int OW, long long ys_n, long long ys_h, long long ys_w, int up, int down, int padx0,
                                 int pady0, cudaStream_t st);

namespace {

struct UpfirdnParams {
    const void* x;
    const float* f;
    void* y;
    int N, C, H, W;
    long long xs_n, xs_c, xs_h, xs_w;
    int fh, fw;
    long long fs_h, fs_w;
    int OH, OW;
    long long ys_n, ys_c, ys_h, ys_w;
    int upx, upy, downx, downy, padx0, pady0, flip;
    float gain;
};

constexpr int MAX_TAPS_SMEM = 32 * 32;

__device__ __forceinline__ int pos_mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

// Stage the (optionally reversed) taps in shared memory, pre-multiplied by gain.  g[ky*fw+kx].
__device__ __forceinline__ void stage_taps(const UpfirdnParams& p, float* sf) {
    for (int t = threadIdx.x; t < p.fh * p.fw; t += blockDim.x) {
        int ky = t / p.fw, kx = t - ky * p.fw;
        int sy = p.flip ? ky : p.fh - 1 - ky;
        int sx = p.flip ? kx : p.fw - 1 - kx;
        sf[t] = p.f[sy * p.fs_h + sx * p.fs_w] * p.gain;
    }
    __syncthreads();
}

// ---- W-contiguous layout -----------------------------------------------------------------------------------------
// grid: x = ceil(OW / 128) * OH tiles folded, y = N*C planes (looped).  UPX.. = 0 means "run-time value".
template <class T, int UPX, int UPY, int DOWNX, int DOWNY, int FW, int FH>
__global__ void __launch_bounds__(128) upfirdn2d_row_kernel(UpfirdnParams p) {
    typedef typename Acc<T>::type S;
    extern __shared__ float sf[];
    stage_taps(p, sf);
    const int upx = UPX ? UPX : p.upx, upy = UPY ? UPY : p.upy;
    const int downx = DOWNX ? DOWNX : p.downx, downy = DOWNY ? DOWNY : p.downy;
    const int fw = FW ? FW : p.fw, fh = FH ? FH : p.fh;
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const int tiles_x = (p.OW + 127) / 128;
    const long long tiles = (long long)tiles_x * p.OH;
    const long long planes = (long long)p.N * p.C;
    for (long long work = blockIdx.x; work < tiles * planes; work += gridDim.x) {
        const long long plane = work / tiles;
        const int tile = (int)(work - plane * tiles);
        const int oy = tile / tiles_x;
        const int ox = (tile - oy * tiles_x) * 128 + threadIdx.x;
        if (ox >= p.OW) continue;
        const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
        const T* xp = x + n * p.xs_n + c * p.xs_c;
        const int by = oy * downy - p.pady0, bx = ox * downx - p.padx0;   // padded-grid origin of the footprint
        const int ky0 = pos_mod(-by, upy), kx0 = pos_mod(-bx, upx);
        S acc = (S)0;
        for (int ky = ky0; ky < fh; ky += upy) {
            const int iy = (by + ky) / upy;
            if (iy < 0 || iy >= p.H) continue;
            const T* xr = xp + iy * p.xs_h;
            const float* fr = sf + ky * fw;
#pragma unroll 4
            for (int kx = kx0; kx < fw; kx += upx) {
                const int ix = (bx + kx) / upx;
                if (ix >= 0 && ix < p.W) acc += to_acc<T>(xr[ix * p.xs_w]) * (S)fr[kx];
            }
        }
        y[n * p.ys_n + c * p.ys_c + oy * p.ys_h + ox * p.ys_w] = from_acc<T>(acc);
    }
}


// ---- small W-contiguous planes (the fp32 4x4..32x32 blocks of G and D, NCHW) ---------------------------------------
// A CTA stages PL whole input planes in shared memory with coalesced loads (a plane is one contiguous run of H*W
// elements), then its threads produce the PL output planes, again as contiguous runs.  Every tap is a shared-memory
// read; global memory sees each input element once.  Same index algebra as the row kernel.
template <class T, int UPX, int UPY, int DOWNX, int DOWNY, int FW, int FH>
__global__ void __launch_bounds__(256) upfirdn2d_plane_kernel(UpfirdnParams p, int PL) {
    typedef typename Acc<T>::type S;
    extern __shared__ float smem_f[];
    const int upx = UPX ? UPX : p.upx, upy = UPY ? UPY : p.upy;
    const int downx = DOWNX ? DOWNX : p.downx, downy = DOWNY ? DOWNY : p.downy;
    const int fw = FW ? FW : p.fw, fh = FH ? FH : p.fh;
    float* sf = smem_f;                       // taps
    S* sx = (S*)(smem_f + ((fh * fw + 3) & ~3));   // PL staged planes
    stage_taps(p, sf);
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const int in_sz = p.H * p.W, out_sz = p.OH * p.OW;
    const long long planes = (long long)p.N * p.C;
    for (long long g0 = (long long)blockIdx.x * PL; g0 < planes; g0 += (long long)gridDim.x * PL) {
        const int npl = (int)((planes - g0) < PL ? (planes - g0) : PL);
        for (int i = threadIdx.x; i < npl * in_sz; i += 256) {
            const int pl = i / in_sz, e = i - pl * in_sz;
            const long long plane = g0 + pl;
            const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
            sx[i] = to_acc<T>(x[n * p.xs_n + c * p.xs_c + e]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < npl * out_sz; i += 256) {
            const int pl = i / out_sz, e = i - pl * out_sz;
            const int oy = e / p.OW, ox = e - oy * p.OW;
            const S* xp = sx + pl * in_sz;
            const int by = oy * downy - p.pady0, bx = ox * downx - p.padx0;
            const int ky0 = pos_mod(-by, upy), kx0 = pos_mod(-bx, upx);
            S acc = (S)0;
            if constexpr (FH != 0) {
                constexpr int NKY = (FH + UPY - 1) / UPY, NKX = (FW + UPX - 1) / UPX;
#pragma unroll
                for (int j = 0; j < NKY; j++) {
                    const int ky = ky0 + j * UPY;
                    const int iy = (by + ky) / UPY;
                    if (ky < FH && iy >= 0 && iy < p.H) {
#pragma unroll
                        for (int i2 = 0; i2 < NKX; i2++) {
                            const int kx = kx0 + i2 * UPX;
                            const int ix = (bx + kx) / UPX;
                            if (kx < FW && ix >= 0 && ix < p.W) acc += xp[iy * p.W + ix] * (S)sf[ky * FW + kx];
                        }
                    }
                }
            } else {
                for (int ky = ky0; ky < fh; ky += upy) {
                    const int iy = (by + ky) / upy;
                    if (iy < 0 || iy >= p.H) continue;
                    for (int kx = kx0; kx < fw; kx += upx) {
                        const int ix = (bx + kx) / upx;
                        if (ix >= 0 && ix < p.W) acc += xp[iy * p.W + ix] * (S)sf[ky * fw + kx];
                    }
                }
            }
            const long long plane = g0 + pl;
            const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
            y[n * p.ys_n + c * p.ys_c + e] = from_acc<T>(acc);
        }
        __syncthreads();
    }
}


// 4x4 taps, up = down = 1, non-negative padding on small W-contiguous planes (the blur around the resampling convolutions of
// the fp32 4x4..16x16 blocks): the planes are staged ZERO-PADDED, so the inner loop has no bounds checks, and a thread
// produces a strip of 4 adjacent outputs from 4 x (16+8+4)-byte shared-memory reads -- ~20 instructions per output instead
// of ~160 for the gather form, which was issue-bound at 0.6 TB/s (ncu, profiles/).
struct Plane44Div {
    FastDiv in_sz, W, strips_per_plane, OW4, C;
};

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_plane44_kernel(UpfirdnParams p, int PL, int PW, int PH, Plane44Div dv) {
    extern __shared__ float smem_f[];
    float* sf = smem_f;                 // 16 taps
    float* sx = smem_f + 16;            // PL padded planes of PH x PW floats (PW a multiple of 4, 16-byte aligned rows)
    stage_taps(p, sf);
    float g[16];
#pragma unroll
    for (int i = 0; i < 16; i++) g[i] = sf[i];
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const int in_sz = p.H * p.W;
    const int OW4 = (p.OW + 3) >> 2;
    const int pplane = PH * PW;
    for (int i = threadIdx.x; i < PL * pplane; i += 256) sx[i] = 0.f;     // borders stay zero for the whole kernel
    __syncthreads();
    const int planes = p.N * p.C;
    const bool vec_out = (p.OW & 3) == 0 && sizeof(T) == 4;
    for (int g0 = blockIdx.x * PL; g0 < planes; g0 += gridDim.x * PL) {
        const int npl = (planes - g0) < PL ? (planes - g0) : PL;
        for (int i = threadIdx.x; i < npl * in_sz; i += 256) {
            uint32_t e, ix;
            const uint32_t pl = fd_divmod((uint32_t)i, dv.in_sz, e);
            const uint32_t iy = fd_divmod(e, dv.W, ix);
            uint32_t c;
            const uint32_t n = fd_divmod((uint32_t)g0 + pl, dv.C, c);
            sx[pl * pplane + (iy + p.pady0) * PW + ix + p.padx0] = (float)to_acc<T>(x[n * p.xs_n + c * p.xs_c + e]);
        }
        __syncthreads();
        const int strips = npl * p.OH * OW4;
        for (int i = threadIdx.x; i < strips; i += 256) {
            uint32_t r, xs;
            const uint32_t pl = fd_divmod((uint32_t)i, dv.strips_per_plane, r);
            const uint32_t oy = fd_divmod(r, dv.OW4, xs);
            const int x0 = (int)xs << 2;
            const float* row = sx + pl * pplane + oy * PW + x0;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int ky = 0; ky < 4; ky++) {
                const float4 u = *reinterpret_cast<const float4*>(row + ky * PW);
                const float2 v = *reinterpret_cast<const float2*>(row + ky * PW + 4);
                const float w = row[ky * PW + 6];
                const float g0_ = g[ky * 4 + 0], g1_ = g[ky * 4 + 1], g2_ = g[ky * 4 + 2], g3_ = g[ky * 4 + 3];
                a0 += u.x * g0_ + u.y * g1_ + u.z * g2_ + u.w * g3_;
                a1 += u.y * g0_ + u.z * g1_ + u.w * g2_ + v.x * g3_;
                a2 += u.z * g0_ + u.w * g1_ + v.x * g2_ + v.y * g3_;
                a3 += u.w * g0_ + v.x * g1_ + v.y * g2_ + w * g3_;
            }
            uint32_t c;
            const uint32_t n = fd_divmod((uint32_t)g0 + pl, dv.C, c);
            T* yp = y + n * p.ys_n + c * p.ys_c + oy * p.OW + x0;
            if (vec_out) {
                *reinterpret_cast<float4*>(yp) = make_float4(a0, a1, a2, a3);
            } else {
                yp[0] = from_acc<T>(a0);
                if (x0 + 1 < p.OW) yp[1] = from_acc<T>(a1);
                if (x0 + 2 < p.OW) yp[2] = from_acc<T>(a2);
                if (x0 + 3 < p.OW) yp[3] = from_acc<T>(a3);
            }
        }
        __syncthreads();
    }
}

// ---- channels-last layout ------------------------------------------------------------------------------------------
// One thread = one output pixel x VEC channels.  Thread order: channel vectors fastest, then ox, oy, n  -> a warp
// reads/writes contiguous 512 bytes whenever C*sizeof(T) >= 512.
struct VecDiv {
    FastDiv cvecs, OW, OH;
};

template <class T, int UPX, int UPY, int DOWNX, int DOWNY, int FW, int FH>
__global__ void __launch_bounds__(256) upfirdn2d_vec_kernel(UpfirdnParams p, VecDiv dv) {
    constexpr int VEC = Vec16<T>::N;
    typedef typename Acc<T>::type S;
    extern __shared__ float sf[];
    stage_taps(p, sf);
    const int upx = UPX ? UPX : p.upx, upy = UPY ? UPY : p.upy;
    const int downx = DOWNX ? DOWNX : p.downx, downy = DOWNY ? DOWNY : p.downy;
    const int fw = FW ? FW : p.fw, fh = FH ? FH : p.fh;
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const int cvecs = p.C / VEC;
    const long long total = (long long)p.N * p.OH * p.OW * cvecs;
    // index split with multiply-shift division: (pixel-vector index within one image) -> cv, ox, oy; images walked by the outer loop
    const uint32_t per_img = (uint32_t)(p.OH * p.OW * cvecs);
    (void)per_img;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)total; i += gridDim.x * blockDim.x) {   // total < 2^31 (launcher)
        uint32_t cvu, oxu, oyu;
        uint32_t r = fd_divmod(i, dv.cvecs, cvu);
        r = fd_divmod(r, dv.OW, oxu);
        const int n = (int)fd_divmod(r, dv.OH, oyu);
        const int cv = (int)cvu, ox = (int)oxu, oy = (int)oyu;
        const T* xp = x + n * p.xs_n + (long long)cv * VEC;        // xs_c == 1
        const int by = oy * downy - p.pady0, bx = ox * downx - p.padx0;
        const int ky0 = pos_mod(-by, upy), kx0 = pos_mod(-bx, upx);
        S acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[k] = (S)0;
        auto tap = [&](int ky, int kx) {
            const int iy = (by + ky) / upy, ix = (bx + kx) / upx;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                const S w = (S)sf[ky * fw + kx];
                Vec16<T> v = ld16(xp + iy * p.xs_h + ix * p.xs_w);
#pragma unroll
                for (int k = 0; k < VEC; k++) acc[k] += to_acc<T>(v.v[k]) * w;
            }
        };
        if constexpr (FH != 0) {
            constexpr int NKY = (FH + UPY - 1) / UPY, NKX = (FW + UPX - 1) / UPX;
#pragma unroll
            for (int j = 0; j < NKY; j++) {
                const int ky = ky0 + j * UPY;
                if (ky < FH) {
#pragma unroll
                    for (int i2 = 0; i2 < NKX; i2++) {
                        const int kx = kx0 + i2 * UPX;
                        if (kx < FW) tap(ky, kx);
                    }
                }
            }
        } else {
            for (int ky = ky0; ky < fh; ky += upy)
                for (int kx = kx0; kx < fw; kx += upx) tap(ky, kx);
        }
        Vec16<T> o;
#pragma unroll
        for (int k = 0; k < VEC; k++) o.v[k] = from_acc<T>(acc[k]);
        st16(y + n * p.ys_n + oy * p.ys_h + ox * p.ys_w + (long long)cv * VEC, o);
    }
}

// ---- fully generic fallback (any strides) ------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(UpfirdnParams p, int c_fastest) {
    typedef typename Acc<T>::type S;
    extern __shared__ float sf[];
    const bool taps_in_smem = p.fh * p.fw <= MAX_TAPS_SMEM;
    if (taps_in_smem) stage_taps(p, sf);
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const long long total = (long long)p.N * p.C * p.OH * p.OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int n, c, oy, ox;
        long long r = i;
        if (c_fastest) { c = (int)(r % p.C); r /= p.C; ox = (int)(r % p.OW); r /= p.OW; oy = (int)(r % p.OH); n = (int)(r / p.OH); }
        else { ox = (int)(r % p.OW); r /= p.OW; oy = (int)(r % p.OH); r /= p.OH; c = (int)(r % p.C); n = (int)(r / p.C); }
        const T* xp = x + n * p.xs_n + c * p.xs_c;
        const int by = oy * p.downy - p.pady0, bx = ox * p.downx - p.padx0;
        const int ky0 = pos_mod(-by, p.upy), kx0 = pos_mod(-bx, p.upx);
        S acc = (S)0;
        for (int ky = ky0; ky < p.fh; ky += p.upy) {
            const int iy = (by + ky) / p.upy;
            if (iy < 0 || iy >= p.H) continue;
            for (int kx = kx0; kx < p.fw; kx += p.upx) {
                const int ix = (bx + kx) / p.upx;
                if (ix < 0 || ix >= p.W) continue;
                float w;
                if (taps_in_smem) w = sf[ky * p.fw + kx];
                else {
                    int sy = p.flip ? ky : p.fh - 1 - ky, sx = p.flip ? kx : p.fw - 1 - kx;
                    w = p.f[sy * p.fs_h + sx * p.fs_w] * p.gain;
                }
                acc += to_acc<T>(xp[iy * p.xs_h + ix * p.xs_w]) * (S)w;
            }
        }
        y[n * p.ys_n + c * p.ys_c + oy * p.ys_h + ox * p.ys_w] = from_acc<T>(acc);
    }
}

template <class T, int UPX, int UPY, int DOWNX, int DOWNY, int FW, int FH>
int launch_row(const UpfirdnParams& p, cudaStream_t st) {
    long long tiles = (long long)((p.OW + 127) / 128) * p.OH * p.N * p.C;
    long long cap = (long long)gt_num_sms() * 16 * 4;
    unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    upfirdn2d_row_kernel<T, UPX, UPY, DOWNX, DOWNY, FW, FH><<<grid, 128, p.fh * p.fw * sizeof(float), st>>>(p);
    GT_CUDA_LAUNCH_CHECK("gt_upfirdn2d(row)");
    return GT_OK;
}


constexpr int PLANE_MAX_ELEMS = 8192;   // staged input elements per CTA (32 KB of fp32)

template <class T, int UPX, int UPY, int DOWNX, int DOWNY, int FW, int FH>
int launch_plane(const UpfirdnParams& p, cudaStream_t st) {
    typedef typename Acc<T>::type S;
    const int in_sz = p.H * p.W;
    const long long planes = (long long)p.N * p.C;
    int PL = PLANE_MAX_ELEMS / in_sz;
    if (PL > 16) PL = 16;
    // keep at least ~2 CTAs per SM busy
    while (PL > 1 && (planes + PL - 1) / PL < (long long)gt_num_sms() * 2) PL >>= 1;
    long long groups = (planes + PL - 1) / PL;
    long long cap = (long long)gt_num_sms() * 8;
    unsigned grid = (unsigned)(groups < cap ? groups : cap);
    size_t smem = (size_t)((p.fh * p.fw + 3) & ~3) * sizeof(float) + (size_t)PL * in_sz * sizeof(S);
    upfirdn2d_plane_kernel<T, UPX, UPY, DOWNX, DOWNY, FW, FH><<<grid, 256, smem, st>>>(p, PL);
    GT_CUDA_LAUNCH_CHECK("gt_upfirdn2d(plane)");
    return GT_OK;
}

template <class T>
int launch_plane44(const UpfirdnParams& p, cudaStream_t st) {
    const int PW = ((p.OW + 3) & ~3) + 4, PH = p.OH + 3;
    const long long planes = (long long)p.N * p.C;
    int PL = PLANE_MAX_ELEMS / (PW * PH);
    if (PL > 16) PL = 16;
    while (PL > 1 && (planes + PL - 1) / PL < (long long)gt_num_sms() * 2) PL >>= 1;
    long long groups = (planes + PL - 1) / PL;
    long long cap = (long long)gt_num_sms() * 6;
    unsigned grid = (unsigned)(groups < cap ? groups : cap);
    size_t smem = (size_t)(16 + PL * PW * PH) * sizeof(float);
    Plane44Div dv;
    dv.in_sz = make_fastdiv((uint32_t)(p.H * p.W));
    dv.W = make_fastdiv((uint32_t)p.W);
    dv.strips_per_plane = make_fastdiv((uint32_t)(p.OH * ((p.OW + 3) >> 2)));
    dv.OW4 = make_fastdiv((uint32_t)((p.OW + 3) >> 2));
    dv.C = make_fastdiv((uint32_t)p.C);
    upfirdn2d_plane44_kernel<T><<<grid, 256, smem, st>>>(p, PL, PW, PH, dv);
    GT_CUDA_LAUNCH_CHECK("gt_upfirdn2d(plane44)");
    return GT_OK;
}

template <class T, int UPX, int UPY, int DOWNX, int DOWNY, int FW, int FH>
int launch_vec(const UpfirdnParams& p, cudaStream_t st) {
    constexpr int VEC = Vec16<T>::N;
    long long total = (long long)p.N * p.OH * p.OW * (p.C / VEC);
    long long blocks = (total + 255) / 256;
    long long cap = (long long)gt_num_sms() * 8 * 4;
    unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
    VecDiv dv;
    dv.cvecs = make_fastdiv((uint32_t)(p.C / VEC));
    dv.OW = make_fastdiv((uint32_t)p.OW);
    dv.OH = make_fastdiv((uint32_t)p.OH);
    upfirdn2d_vec_kernel<T, UPX, UPY, DOWNX, DOWNY, FW, FH><<<grid, 256, p.fh * p.fw * sizeof(float), st>>>(p, dv);
    GT_CUDA_LAUNCH_CHECK("gt_upfirdn2d(vec)");
    return GT_OK;
}

#define GT_UPFIRDN_CASE(LAUNCH, UX, UY, DX, DY, FW_, FH_)                                                                     \
    if (p.upx == UX && p.upy == UY && p.downx == DX && p.downy == DY && p.fw == FW_ && p.fh == FH_)                             \
        return LAUNCH<T, UX, UY, DX, DY, FW_, FH_>(p, st);

template <class T>
int dispatch(const UpfirdnParams& p, cudaStream_t st) {
    constexpr int VEC = Vec16<T>::N;
    const bool small_f = p.fh * p.fw <= MAX_TAPS_SMEM;
    const bool al = ((((uintptr_t)p.x) | ((uintptr_t)p.y)) & 15) == 0;
    const bool cl = small_f && al && p.xs_c == 1 && p.ys_c == 1 && p.C % VEC == 0 && p.C >= VEC && (long long)p.N * p.OH * p.OW * (p.C / VEC) < (1ll << 31) && p.xs_w % VEC == 0 && p.xs_h % VEC == 0 &&
                    p.xs_n % VEC == 0 && p.ys_w % VEC == 0 && p.ys_h % VEC == 0 && p.ys_n % VEC == 0;
    if (cl) {
        GT_UPFIRDN_CASE(launch_vec, 1, 1, 1, 1, 4, 4)
        GT_UPFIRDN_CASE(launch_vec, 2, 2, 1, 1, 4, 4)
        GT_UPFIRDN_CASE(launch_vec, 1, 1, 2, 2, 4, 4)
        return launch_vec<T, 0, 0, 0, 0, 0, 0>(p, st);
    }
    if (small_f && p.xs_w == 1 && p.ys_w == 1 && p.xs_h == p.W && p.ys_h == p.OW && p.H * p.W <= PLANE_MAX_ELEMS && p.OH * p.OW <= 4 * PLANE_MAX_ELEMS &&
        (long long)p.N * p.C >= 64 && (long long)p.N * p.C < (1ll << 30) && sizeof(T) <= 4) {
        if (p.upx == 1 && p.upy == 1 && p.downx == 1 && p.downy == 1 && p.fw == 4 && p.fh == 4 && p.padx0 >= 0 && p.pady0 >= 0 &&
            p.OW + 3 - p.W - p.padx0 >= 0 && p.OH + 3 - p.H - p.pady0 >= 0 && p.ys_h == p.OW &&
            (((p.OW + 3) & ~3) + 4) * (p.OH + 3) <= PLANE_MAX_ELEMS && (p.ys_n % 4 == 0) && (p.ys_c % 4 == 0) && (((uintptr_t)p.y) & 15) == 0)
            return launch_plane44<T>(p, st);
        GT_UPFIRDN_CASE(launch_plane, 1, 1, 1, 1, 4, 4)
        GT_UPFIRDN_CASE(launch_plane, 2, 2, 1, 1, 4, 4)
        GT_UPFIRDN_CASE(launch_plane, 1, 1, 2, 2, 4, 4)
        if (p.fh * p.fw <= 64) return launch_plane<T, 0, 0, 0, 0, 0, 0>(p, st);
    }
    if (small_f && p.xs_w == 1 && p.ys_w == 1) {
        GT_UPFIRDN_CASE(launch_row, 1, 1, 1, 1, 4, 4)
        GT_UPFIRDN_CASE(launch_row, 2, 2, 1, 1, 4, 4)
        GT_UPFIRDN_CASE(launch_row, 1, 1, 2, 2, 4, 4)
        GT_UPFIRDN_CASE(launch_row, 2, 1, 1, 1, 12, 1)
        GT_UPFIRDN_CASE(launch_row, 1, 2, 1, 1, 1, 12)
        GT_UPFIRDN_CASE(launch_row, 1, 1, 2, 1, 12, 1)
        GT_UPFIRDN_CASE(launch_row, 1, 1, 1, 2, 1, 12)
        return launch_row<T, 0, 0, 0, 0, 0, 0>(p, st);
    }
    long long total = (long long)p.N * p.C * p.OH * p.OW;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)gt_num_sms() * 32;
    unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
    size_t smem = small_f ? p.fh * p.fw * sizeof(float) : 0;
    upfirdn2d_generic_kernel<T><<<grid, 256, smem, st>>>(p, p.xs_c == 1 ? 1 : 0);
    GT_CUDA_LAUNCH_CHECK("gt_upfirdn2d(generic)");
    return GT_OK;
}

}  // namespace

extern "C" int gt_upfirdn2d(const void* x, const float* f, void* y, int dtype, int N, int C, int H, int W, long long xs_n, long long xs_c,
                            long long xs_h, long long xs_w, int fh, int fw, long long fs_h, long long fs_w, int OH, int OW, long long ys_n,
                            long long ys_c, long long ys_h, long long ys_w, int upx, int upy, int downx, int downy, int padx0, int pady0,
                            int flip, float gain, void* stream) {
    GT_REQUIRE(x && f && y, "gt_upfirdn2d: null pointer");
    GT_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "gt_upfirdn2d: x has zero size");
    GT_REQUIRE(fh >= 1 && fw >= 1, "gt_upfirdn2d: f must be at least 1x1");
    GT_REQUIRE(upx >= 1 && upy >= 1, "gt_upfirdn2d: upsampling factor must be at least 1");
    GT_REQUIRE(downx >= 1 && downy >= 1, "gt_upfirdn2d: downsampling factor must be at least 1");
    GT_REQUIRE(OH >= 1 && OW >= 1, "gt_upfirdn2d: output must be at least 1x1");
    UpfirdnParams p;
    p.x = x; p.f = f; p.y = y; p.N = N; p.C = C; p.H = H; p.W = W;
    p.xs_n = xs_n; p.xs_c = xs_c; p.xs_h = xs_h; p.xs_w = xs_w;
    p.fh = fh; p.fw = fw; p.fs_h = fs_h; p.fs_w = fs_w; p.OH = OH; p.OW = OW;
    p.ys_n = ys_n; p.ys_c = ys_c; p.ys_h = ys_h; p.ys_w = ys_w;
    p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0; p.flip = flip ? 1 : 0; p.gain = gain;
    cudaStream_t st = (cudaStream_t)stream;
    if (upx == 1 && upy == 1 && downx == 1 && downy == 1 && fh == 4 && fw == 4 && xs_c == 1 && ys_c == 1 && (dtype == GT_F16 || dtype == GT_F32) &&
        (long long)N * OH * OW * C >= (1 << 20)) {
        // the blur around the resampling convolutions on channels-last tensors: TMA-staged kernel (upfirdn2d_tma.cu)
        const int rc = gt_upfirdn2d_try_tma(x, f, fs_h, fs_w, flip ? 1 : 0, gain, y, dtype, N, C, H, W, xs_n, xs_h, xs_w, OH, OW, ys_n, ys_h, ys_w, padx0, pady0, st);
        if (rc >= 0) return rc;
    }
    if (upx == upy && downx == downy && ((upx == 2 && downx == 1) || (upx == 1 && downx == 2)) && fh == 4 && fw == 4 && xs_c == 1 && ys_c == 1 &&
        (dtype == GT_F16 || dtype == GT_F32) && (long long)N * C * (upx == 2 ? (long long)OH * OW : (long long)H * W) >= (1 << 20) && gt_stream_variant() == 0) {
        // factor-2 resampling of the skip branches on channels-last tensors (upfirdn2d_tma_strided.cu)
        const int rc = gt_upfirdn2d_try_tma_strided(x, f, fs_h, fs_w, flip ? 1 : 0, gain, y, dtype, N, C, H, W, xs_n, xs_h, xs_w, OH, OW, ys_n, ys_h, ys_w, upx, downx, padx0, pady0, st);
        if (rc >= 0) return rc;
    }
    switch (dtype) {
        case GT_F32: return dispatch<float>(p, st);
        case GT_F16: return dispatch<__half>(p, st);
        case GT_F64: return dispatch<double>(p, st);
    }
    gt_set_error("gt_upfirdn2d: unsupported dtype code %d", dtype);
    return GT_ERR_ARG;
}
