// modulated.cu -- the element-wise halves of the training-mode modulated convolution on channels-last tensors, fused.
//
// Reference (S3/training/networks_stylegan2.py:68-77 and :325-327, S3 = /root/reference/src/models/stylegan3):
//     x = x * styles.to(x.dtype)[:, :, None, None]                       one full pass (+ 3 in backward)
//     x = conv2d_resample(x, w, ...)
//     x = fma(x, dcoefs.to(x.dtype)[:, :, None, None], noise)            one full pass (+ 3-4 in backward, OPS/fma.py:15-58)
//     x = bias_act(x, b, act, gain, clamp)                                one full pass (+ 1 in backward + db reduction)
// Here:
//     gt_mod_scale_fwd        y = x * s[n,c]
//     gt_mod_scale_bwd        gx = gy * s[n,c]   and   gs[n,c] = sum_hw gy * x          one pass, deterministic reduction
//     gt_demod_act_fwd        y = clamp(act(x * d[n,c] + noise[n,hw] + b[c]) * gain)     one pass
//     gt_demod_act_bwd        g1 = gy * gain * act'(y) * [|y| < clamp];  gx = g1 * d[n,c];
//                             gd[n,c] = sum_hw g1 * x;  s0[n,c] = sum_hw g1 (-> db);  gnoise[n,hw] = sum_c g1      one pass
// All tensors are [N, P, C] (P = H*W pixels, C contiguous, C % (16 / sizeof(T)) == 0); s, d, gs, gd, s0 are fp32 [N,C].
// Math is fp32 for fp16 I/O (as OPS/bias_act.cu:15-18).  act: 1 = linear, 3 = lrelu (the two on the StyleGAN2 path).
// Reductions over pixels: a CTA owns (n, band of pixels); per-CTA partials go to a workspace and are summed in band
// order by a second kernel, so results are bit-reproducible.
#include "gt_common.cuh"
#include "stream_bulk.cuh"
#include "hot_act.cuh"

namespace {

constexpr int A_LINEAR = 1, A_LRELU = 3;
constexpr int MAXJ = 4;   // channel-vector chunks per thread: C <= 32 lanes * MAXJ * VEC

template <class T>
__global__ void __launch_bounds__(256) mod_scale_fwd_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ y, int C, long long P,
                                                            long long total_vecs) {
    constexpr int VEC = Vec16<T>::N;
    const int cvecs = C / VEC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vecs; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % cvecs);
        const long long n = (i / cvecs) / P;
        const Vec16<T> v = ld16_stream(x + i * VEC);
        const float* sp = s + n * C + cv * VEC;
        Vec16<T> o;
#pragma unroll
        for (int k = 0; k < VEC; k++) o.v[k] = from_acc<T>((float)to_acc<T>(v.v[k]) * (float)to_acc<T>(from_acc<T>(sp[k])));
        st16_stream(y + i * VEC, o);
    }
}

// Shared structure of the two backward kernels: grid = (bands, N); lanes = min(cvecs, 32) threads cover the channel
// vectors of one pixel (each thread owns cv = lane + lanes * j), 256 / lanes pixel rows are in flight per CTA.
struct BandGeom {
    int cvecs, lanes, nj, rgroups, rg, lane;
};
template <int VEC>
__device__ __forceinline__ BandGeom band_geom(int C) {
    BandGeom g;
    g.cvecs = C / VEC;
    g.lanes = g.cvecs < 32 ? g.cvecs : 32;
    g.nj = (g.cvecs + g.lanes - 1) / g.lanes;
    g.rgroups = 256 / g.lanes;
    g.rg = threadIdx.x / g.lanes;
    g.lane = threadIdx.x - g.rg * g.lanes;
    return g;
}

// Combine the per-thread accumulators of the row groups in a fixed order and write this CTA's partial [C] vector.
template <int VEC>
__device__ __forceinline__ void write_partial(const BandGeom& g, float (&acc)[MAXJ][VEC], float* red, float* __restrict__ out, int C) {
#pragma unroll
    for (int j = 0; j < MAXJ; j++) {
        if (j < g.nj) {   // uniform across the CTA
            __syncthreads();
#pragma unroll
            for (int k = 0; k < VEC; k++) red[threadIdx.x * VEC + k] = acc[j][k];
            __syncthreads();
            const int cv = g.lane + g.lanes * j;
            if (g.rg == 0 && cv < g.cvecs) {
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    float sum = 0.f;
                    for (int r = 0; r < g.rgroups; r++) sum += red[(r * g.lanes + g.lane) * VEC + k];
                    out[cv * VEC + k] = sum;
                }
            }
        }
    }
}

template <class T>
__global__ void __launch_bounds__(256) mod_scale_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x, const float* __restrict__ s,
                                                            T* __restrict__ gx, float* __restrict__ partial, int C, long long P) {
    constexpr int VEC = Vec16<T>::N;
    __shared__ float red[256 * VEC];
    const BandGeom g = band_geom<VEC>(C);
    const int band = blockIdx.x, bands = gridDim.x, n = blockIdx.y;
    float acc[MAXJ][VEC];
#pragma unroll
    for (int j = 0; j < MAXJ; j++)
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[j][k] = 0.f;
    const long long base = (long long)n * P * C;
    for (long long p = band + (long long)bands * g.rg; p < P; p += (long long)bands * g.rgroups) {
#pragma unroll
        for (int j = 0; j < MAXJ; j++) {
            const int cv = g.lane + g.lanes * j;
            if (j < g.nj && cv < g.cvecs) {
                const long long e = base + p * C + (long long)cv * VEC;
                const Vec16<T> gv = ld16_stream(gy + e), xv = ld16_stream(x + e);
                const float* sp = s + (long long)n * C + cv * VEC;
                Vec16<T> o;
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    const float gg = (float)to_acc<T>(gv.v[k]);
                    o.v[k] = from_acc<T>(gg * (float)to_acc<T>(from_acc<T>(sp[k])));
                    acc[j][k] += gg * (float)to_acc<T>(xv.v[k]);
                }
                st16_stream(gx + e, o);
            }
        }
    }
    write_partial<VEC>(g, acc, red, partial + ((long long)n * bands + band) * C, C);
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float u, float alpha, float gain, float clampv) {
    float v = u;
    if (ACT == A_LRELU) v = u > 0.f ? u : u * alpha;
    v *= gain;
    if (clampv >= 0.f) v = fminf(fmaxf(v, -clampv), clampv);
    return v;
}
// derivative factor decided from the saved OUTPUT, as the reference plugin does (OPS/bias_act.cu:132-142)
template <int ACT>
__device__ __forceinline__ float act_slope(float y, float alpha, float gain, float clampv) {
    float m = gain;
    if (ACT == A_LRELU) m = (y > 0.f) ? gain : gain * alpha;
    if (clampv >= 0.f && !(fabsf(y) < clampv)) m = 0.f;
    return m;
}

template <class T, int ACT>
__global__ void __launch_bounds__(256) demod_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ d, const T* __restrict__ nz,
                                                            const T* __restrict__ b, T* __restrict__ y, int C, long long P, long long total_vecs,
                                                            float alpha, float gain, float clampv) {
    constexpr int VEC = Vec16<T>::N;
    const int cvecs = C / VEC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vecs; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % cvecs);
        const long long pix = i / cvecs;       // n * P + p
        const long long n = pix / P;
        const Vec16<T> v = ld16_stream(x + i * VEC);
        const float noise = nz ? (float)to_acc<T>(nz[pix]) : 0.f;
        Vec16<T> bv;
        if (b) bv = ld16(b + cv * VEC);
        const float* dp = d ? d + n * C + cv * VEC : nullptr;
        Vec16<T> o;
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            float u = (float)to_acc<T>(v.v[k]);
            if (dp) u *= (float)to_acc<T>(from_acc<T>(dp[k]));
            u += noise;
            // the reference materialises fma(x, d, noise) in T before bias_act reads it back (networks_stylegan2.py:71-72);
            // rounding at the same point keeps the sign of near-zero pre-activations -- and with it the lrelu slope -- identical
            if (dp || nz) u = (float)to_acc<T>(from_acc<T>(u));
            if (b) u += (float)to_acc<T>(bv.v[k]);
            o.v[k] = from_acc<T>(act_fwd<ACT>(u, alpha, gain, clampv));
        }
        st16_stream(y + i * VEC, o);
    }
}

template <class T, int ACT>
__global__ void __launch_bounds__(256) demod_act_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ yref, const T* __restrict__ x,
                                                            const float* __restrict__ d, T* __restrict__ gx, T* __restrict__ gnz,
                                                            float* __restrict__ partial1, float* __restrict__ partial0, int C, long long P,
                                                            float alpha, float gain, float clampv) {
    constexpr int VEC = Vec16<T>::N;
    __shared__ float red[256 * VEC];
    const BandGeom g = band_geom<VEC>(C);
    const int band = blockIdx.x, bands = gridDim.x, n = blockIdx.y;
    float acc1[MAXJ][VEC], acc0[MAXJ][VEC];
#pragma unroll
    for (int j = 0; j < MAXJ; j++)
#pragma unroll
        for (int k = 0; k < VEC; k++) acc1[j][k] = acc0[j][k] = 0.f;
    const long long base = (long long)n * P * C;
    // every thread of a row group iterates the same number of times so that the shuffles below stay converged
    const long long iters = (P - band + (long long)bands * g.rgroups - 1) / ((long long)bands * g.rgroups);
    for (long long it = 0; it < iters; it++) {
        const long long p = band + (long long)bands * (g.rg + g.rgroups * it);
        const bool live = p < P;
        float pix_sum = 0.f;
#pragma unroll
        for (int j = 0; j < MAXJ; j++) {
            const int cv = g.lane + g.lanes * j;
            if (live && j < g.nj && cv < g.cvecs) {
                const long long e = base + p * C + (long long)cv * VEC;
                const Vec16<T> gv = ld16_stream(gy + e), yv = ld16_stream(yref + e);
                Vec16<T> xv;
                if (d) xv = ld16_stream(x + e);
                const float* dp = d ? d + (long long)n * C + cv * VEC : nullptr;
                Vec16<T> o;
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    const float g1 = (float)to_acc<T>(gv.v[k]) * act_slope<ACT>((float)to_acc<T>(yv.v[k]), alpha, gain, clampv);
                    const float g1r = (float)to_acc<T>(from_acc<T>(g1));       // the reference materialises g1 in T before the fma backward
                    pix_sum += g1r;
                    acc0[j][k] += g1r;
                    if (dp) {
                        acc1[j][k] += g1r * (float)to_acc<T>(xv.v[k]);
                        o.v[k] = from_acc<T>(g1r * (float)to_acc<T>(from_acc<T>(dp[k])));
                    } else {
                        o.v[k] = from_acc<T>(g1);
                    }
                }
                st16_stream(gx + e, o);
            }
        }
        if (gnz) {
            // sum over the channels of this pixel: lanes of the row group are contiguous lanes of one warp (lanes <= 32)
            for (int off = g.lanes >> 1; off > 0; off >>= 1) pix_sum += __shfl_xor_sync(0xffffffffu, pix_sum, off);
            if (live && g.lane == 0) gnz[(long long)n * P + p] = from_acc<T>(pix_sum);
        }
    }
    if (partial1) write_partial<VEC>(g, acc1, red, partial1 + ((long long)n * bands + band) * C, C);
    write_partial<VEC>(g, acc0, red, partial0 + ((long long)n * bands + band) * C, C);
}


// =====================================================================================================================
// Bulk-staged variants (stream_bulk.cuh): grid = (bands, N); a CTA streams chunks band, band + bands, ... of sample n.
// C divides 256 * VEC, so a thread always sees the same VEC channels: s / d / b live in registers as fp32 pairs, the
// per-channel reductions are register accumulators combined once at the end (fixed order).  Same arithmetic, same
// rounding points as the direct kernels above.
// =====================================================================================================================
template <class T> __device__ __forceinline__ float2 round_io(float2 v);          // round to the I/O type and back
template <> __device__ __forceinline__ float2 round_io<__half>(float2 v) { return __half22float2(__float22half2_rn(v)); }
template <> __device__ __forceinline__ float2 round_io<float>(float2 v) { return v; }

// this thread's VEC channels of a per-(n, c) fp32 vector, rounded to T like `v.to(x.dtype)` in the reference
template <class T>
__device__ __forceinline__ void load_nc(const float* v, int n, int C, float2 (&out)[hot::Lanes<T>::NP]) {
    constexpr int VEC = Vec16<T>::N;
    const float* p = v + (long long)n * C + ((int)threadIdx.x * VEC) % C;
#pragma unroll
    for (int i = 0; i < hot::Lanes<T>::NP; i++) out[i] = round_io<T>(make_float2(p[2 * i], p[2 * i + 1]));
}

// combine the consumers' register accumulators (threads with equal tid % cvecs own the same channels) -> out[C]
template <int VEC>
__device__ __forceinline__ void bulk_write_partial(const float (&acc)[VEC], float* red, float* __restrict__ out, int C) {
    __syncthreads();
    if (threadIdx.x < 256) {
#pragma unroll
        for (int k = 0; k < VEC; k++) red[threadIdx.x * VEC + k] = acc[k];
    }
    __syncthreads();
    const int cvecs = C / VEC;
    if ((int)threadIdx.x < cvecs) {
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            float sum = 0.f;
            for (int g2 = threadIdx.x; g2 < 256; g2 += cvecs) sum += red[g2 * VEC + k];
            out[threadIdx.x * VEC + k] = sum;
        }
    }
}

template <class T>
__global__ void __launch_bounds__(streamk::NTHREADS) mod_scale_fwd_bulk_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ y,
                                                                               int C, long long P) {
    typedef hot::Lanes<T> L;
    extern __shared__ uint8_t smem_raw[];
    const int n = blockIdx.y;
    float2 sv[L::NP];
    load_nc<T>(s, n, C, sv);
    const long long base = (long long)n * P * C;
    streamk::run<T, 1, 6, 16384>(x + base, (const T*)nullptr, (const T*)nullptr, y + base, P * C, smem_raw,
                                 [&](long long, Vec16<T>& v0, const Vec16<T>&, const Vec16<T>&, bool live, uint32_t, int) {
                                     if (!live) return;
#pragma unroll
                                     for (int i = 0; i < L::NP; i++) L::set(v0, i, __fmul2_rn(L::get(v0, i), sv[i]));
                                 });
}

template <class T>
__global__ void __launch_bounds__(streamk::NTHREADS) mod_scale_bwd_bulk_kernel(const T* __restrict__ gy, const T* __restrict__ x, const float* __restrict__ s,
                                                                               T* __restrict__ gx, float* __restrict__ partial, int C, long long P) {
    typedef hot::Lanes<T> L;
    constexpr int VEC = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    __shared__ float red[256 * VEC];
    const int n = blockIdx.y;
    float2 sv[L::NP];
    load_nc<T>(s, n, C, sv);
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) acc[k] = 0.f;
    const long long base = (long long)n * P * C;
    streamk::run<T, 2, 6, 8192>(gy + base, x + base, (const T*)nullptr, gx + base, P * C, smem_raw,
                                [&](long long, Vec16<T>& v0, const Vec16<T>& v1, const Vec16<T>&, bool live, uint32_t, int) {
                                    if (!live) return;
#pragma unroll
                                    for (int i = 0; i < L::NP; i++) {
                                        const float2 gg = L::get(v0, i), xx = L::get(v1, i);
                                        const float2 a = __ffma2_rn(gg, xx, make_float2(acc[2 * i], acc[2 * i + 1]));
                                        acc[2 * i] = a.x;
                                        acc[2 * i + 1] = a.y;
                                        L::set(v0, i, __fmul2_rn(gg, sv[i]));
                                    }
                                });
    bulk_write_partial<VEC>(acc, red, partial + ((long long)n * gridDim.x + blockIdx.x) * C, C);
}

template <class T, int ACT>
__global__ void __launch_bounds__(streamk::NTHREADS) demod_act_fwd_bulk_kernel(const T* __restrict__ x, const float* __restrict__ d, const T* __restrict__ nz,
                                                                               const T* __restrict__ b, T* __restrict__ y, int C, int logC, long long P,
                                                                               float alpha, float gain, float clampv) {
    typedef hot::Lanes<T> L;
    constexpr int VEC = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    const int n = blockIdx.y;
    float2 dv[L::NP], bv[L::NP];
#pragma unroll
    for (int i = 0; i < L::NP; i++) dv[i] = make_float2(1.f, 1.f), bv[i] = make_float2(0.f, 0.f);
    if (d) load_nc<T>(d, n, C, dv);
    if (b) {
        const Vec16<T> braw = ld16(b + ((int)threadIdx.x * VEC) % C);
#pragma unroll
        for (int i = 0; i < L::NP; i++) bv[i] = L::get(braw, i);
    }
    const hot::Params hp = hot::make_params(alpha, gain, clampv);
    const bool clamp_on = clampv >= 0.f;
    const bool pre = (d != nullptr) || (nz != nullptr);
    const T* nzn = nz ? nz + (long long)n * P : nullptr;
    const long long base = (long long)n * P * C;
    streamk::run<T, 1, 6, 16384>(x + base, (const T*)nullptr, (const T*)nullptr, y + base, P * C, smem_raw,
                                 [&](long long, Vec16<T>& v0, const Vec16<T>&, const Vec16<T>&, bool live, uint32_t side_sa, int vo) {
                                     if (!live) return;
                                     float nzv = 0.f;
                                     if (nzn) {       // this chunk's noise values were staged with it: pixel = byte offset / (C * sizeof(T))
                                         T nt;
                                         const uint32_t a = side_sa + (uint32_t)(((vo / (int)sizeof(T)) >> logC) * (int)sizeof(T));
                                         if (sizeof(T) == 2) {
                                             unsigned short u;
                                             asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(a));
                                             nt = *reinterpret_cast<T*>(&u);
                                         } else {
                                             uint32_t u;
                                             asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(a));
                                             nt = *reinterpret_cast<T*>(&u);
                                         }
                                         nzv = (float)to_acc<T>(nt);
                                     }
#pragma unroll
                                     for (int i = 0; i < L::NP; i++) {
                                         float2 u = L::get(v0, i);
                                         // the reference materialises fma(x, d, noise) in T before bias_act reads it back
                                         // (networks_stylegan2.py:71-72): round at the same point
                                         if (pre) u = round_io<T>(__ffma2_rn(u, dv[i], make_float2(nzv, nzv)));
                                         u = __fadd2_rn(u, bv[i]);
                                         L::set(v0, i, clamp_on ? hot::fwd<ACT, true>(u, hp) : hot::fwd<ACT, false>(u, hp));
                                     }
                                 },
                                 nzn, nzn ? (int)((16384 >> logC)) : 0);      // (16384 / (C * sizeof(T))) pixels * sizeof(T) bytes per chunk
}

template <class T, int ACT, bool HAS_D>
__global__ void __launch_bounds__(streamk::NTHREADS) demod_act_bwd_bulk_kernel(const T* __restrict__ gy, const T* __restrict__ yref, const T* __restrict__ x,
                                                                               const float* __restrict__ d, T* __restrict__ gx, T* __restrict__ gnz,
                                                                               float* __restrict__ partial1, float* __restrict__ partial0, int C, int logC,
                                                                               long long P, float alpha, float gain, float clampv) {
    typedef hot::Lanes<T> L;
    constexpr int VEC = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    __shared__ float red[256 * VEC];
    const int n = blockIdx.y;
    float2 dv[L::NP];
#pragma unroll
    for (int i = 0; i < L::NP; i++) dv[i] = make_float2(1.f, 1.f);
    if (HAS_D) load_nc<T>(d, n, C, dv);
    float acc1[VEC], acc0[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) acc1[k] = acc0[k] = 0.f;
    const hot::Params hp = hot::make_params(alpha, gain, clampv);
    const bool clamp_on = clampv >= 0.f;
    const int cvecs = C / VEC;                 // <= 32 and a power of two when gnz is requested (launcher)
    T* gnzn = gnz ? gnz + (long long)n * P : nullptr;
    const long long base = (long long)n * P * C;
    auto body = [&](long long e0, Vec16<T>& v0, const Vec16<T>& v1, const Vec16<T>& v2, bool live, uint32_t, int) {
        float pix_sum = 0.f;
        if (live) {
#pragma unroll
            for (int i = 0; i < L::NP; i++) {
                // g1 = gy * gain * act'(y) * [|y| < clamp], materialised in T before the fma backward like the reference
                const float2 g1 = clamp_on ? hot::bwd<ACT, true>(L::get(v0, i), L::get(v1, i), hp) : hot::bwd<ACT, false>(L::get(v0, i), L::get(v1, i), hp);
                const float2 g1r = round_io<T>(g1);
                pix_sum += g1r.x + g1r.y;
                acc0[2 * i] += g1r.x;
                acc0[2 * i + 1] += g1r.y;
                if (HAS_D) {
                    const float2 a = __ffma2_rn(g1r, L::get(v2, i), make_float2(acc1[2 * i], acc1[2 * i + 1]));
                    acc1[2 * i] = a.x;
                    acc1[2 * i + 1] = a.y;
                    L::set(v0, i, __fmul2_rn(g1r, dv[i]));
                } else {
                    L::set(v0, i, g1);
                }
            }
        }
        if (gnzn) {     // sum over the channels of this pixel: its cvecs vectors sit on adjacent lanes of one warp
            for (int off = cvecs >> 1; off > 0; off >>= 1) pix_sum += __shfl_xor_sync(0xffffffffu, pix_sum, off);
            if (live && ((int)threadIdx.x & (cvecs - 1)) == 0) gnzn[e0 >> logC] = from_acc<T>(pix_sum);
        }
    };
    if (HAS_D) streamk::run<T, 3, 4, 8192>(gy + base, yref + base, x + base, gx + base, P * C, smem_raw, body);
    else streamk::run<T, 2, 6, 8192>(gy + base, yref + base, (const T*)nullptr, gx + base, P * C, smem_raw, body);
    if (HAS_D) bulk_write_partial<VEC>(acc1, red, partial1 + ((long long)n * gridDim.x + blockIdx.x) * C, C);
    bulk_write_partial<VEC>(acc0, red, partial0 + ((long long)n * gridDim.x + blockIdx.x) * C, C);
}

template <class K>
int set_smem(K kernel, int bytes, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        gt_set_error("%s: cannot reserve %d bytes of shared memory: %s", name, bytes, cudaGetErrorString(e));
        return GT_ERR_CUDA;
    }
    return GT_OK;
}

// The bulk path applies when a sample is at least 1 MB, its channel count divides 256 * VEC (a thread's channels are
// then fixed) and everything is 16-byte aligned.
template <class T>
bool bulk_ok(int N, long long P, int C, const void* a, const void* b2, const void* c2, const void* d2) {
    constexpr int VEC = Vec16<T>::N;
    if (gt_stream_variant() != 0) return false;
    if (C % VEC != 0 || (256 * VEC) % C != 0 || (C & (C - 1)) != 0) return false;
    if (P * C * (long long)sizeof(T) < (1ll << 20) || N > 65535) return false;
    return ((((uintptr_t)a) | ((uintptr_t)b2) | ((uintptr_t)c2) | ((uintptr_t)d2)) & 15) == 0;
}
int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}
// CTAs per sample so that the whole grid is one resident wave (2 CTAs per SM)
int bulk_bands(int N, long long P, int C, int esz, int ch_bytes, int max_bands) {
    long long bands = ((long long)gt_num_sms() * 2) / N;
    const long long chunks = (P * C * esz + ch_bytes - 1) / ch_bytes;
    if (bands > chunks) bands = chunks;
    if (bands > max_bands) bands = max_bands;
    if (bands < 1) bands = 1;
    return (int)bands;
}


// =====================================================================================================================
// Second-order passes (the path-length regulariser differentiates the generator's backward, S3/training/loss.py:85-100):
// the derivatives of gt_mod_scale_bwd / gt_demod_act_bwd with respect to their inputs, each as ONE pass instead of the
// ~10 element-wise torch passes the formulas take when written with tensor ops.  Same band / row-group structure as the
// first-order direct kernels; per-(n, c) reductions go through the same partial-sum workspace.
//   mod_scale:   d_gy = ggx * s + ggs * x        d_x = ggs * gy                 d_s[n,c] = sum_hw ggx * gy
//   demod_act:   m = gain * act'(y) * [|y| < clamp],  g1 = gy * m
//                d_gy = m * (ggx * d + ggd * x + ggnz + ggs0)      d_x = ggd * g1      d_d[n,c] = sum_hw ggx * g1
// Any of ggx / ggs / ggd / ggnz / ggs0 may be NULL (that cotangent is absent); outputs that are not wanted are NULL.
// =====================================================================================================================
template <class T>
__global__ void __launch_bounds__(256) mod_scale_bwd2_kernel(const T* __restrict__ ggx, const float* __restrict__ ggs, const T* __restrict__ gy,
                                                             const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ d_gy,
                                                             T* __restrict__ d_x, float* __restrict__ partial, int C, long long P) {
    constexpr int VEC = Vec16<T>::N;
    __shared__ float red[256 * VEC];
    const BandGeom g = band_geom<VEC>(C);
    const int band = blockIdx.x, bands = gridDim.x, n = blockIdx.y;
    float acc[MAXJ][VEC];
#pragma unroll
    for (int j = 0; j < MAXJ; j++)
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[j][k] = 0.f;
    const long long base = (long long)n * P * C;
    for (long long p = band + (long long)bands * g.rg; p < P; p += (long long)bands * g.rgroups) {
#pragma unroll
        for (int j = 0; j < MAXJ; j++) {
            const int cv = g.lane + g.lanes * j;
            if (j < g.nj && cv < g.cvecs) {
                const long long e = base + p * C + (long long)cv * VEC;
                Vec16<T> av, gv, xv, o1, o2;
                if (ggx) av = ld16_stream(ggx + e);
                if (ggx || d_x) gv = ld16_stream(gy + e);
                if (ggs && d_gy) xv = ld16_stream(x + e);
                const float* sp = s + (long long)n * C + cv * VEC;
                const float* gsp = ggs ? ggs + (long long)n * C + cv * VEC : nullptr;
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    const float a = ggx ? (float)to_acc<T>(av.v[k]) : 0.f;
                    const float gsv = gsp ? (float)to_acc<T>(from_acc<T>(gsp[k])) : 0.f;
                    float t = 0.f;
                    if (ggx) t += a * (float)to_acc<T>(from_acc<T>(sp[k]));
                    if (gsp && d_gy) t += gsv * (float)to_acc<T>(xv.v[k]);
                    o1.v[k] = from_acc<T>(t);
                    if (d_x) o2.v[k] = from_acc<T>(gsv * (float)to_acc<T>(gv.v[k]));
                    if (ggx) acc[j][k] += a * (float)to_acc<T>(gv.v[k]);
                }
                if (d_gy) st16_stream(d_gy + e, o1);
                if (d_x) st16_stream(d_x + e, o2);
            }
        }
    }
    if (partial) write_partial<VEC>(g, acc, red, partial + ((long long)n * bands + band) * C, C);
}

template <class T, int ACT>
__global__ void __launch_bounds__(256) demod_act_bwd2_kernel(const T* __restrict__ ggx, const float* __restrict__ ggd, const T* __restrict__ ggnz,
                                                             const float* __restrict__ ggs0, const T* __restrict__ gy, const T* __restrict__ yref,
                                                             const T* __restrict__ x, const float* __restrict__ d, T* __restrict__ d_gy,
                                                             T* __restrict__ d_x, float* __restrict__ partial, int C, long long P, float alpha,
                                                             float gain, float clampv) {
    constexpr int VEC = Vec16<T>::N;
    __shared__ float red[256 * VEC];
    const BandGeom g = band_geom<VEC>(C);
    const int band = blockIdx.x, bands = gridDim.x, n = blockIdx.y;
    float acc[MAXJ][VEC];
#pragma unroll
    for (int j = 0; j < MAXJ; j++)
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[j][k] = 0.f;
    const long long base = (long long)n * P * C;
    const bool need_x = (ggd != nullptr) && (d_gy != nullptr);
    for (long long p = band + (long long)bands * g.rg; p < P; p += (long long)bands * g.rgroups) {
        const float nzv = ggnz ? (float)to_acc<T>(ggnz[(long long)n * P + p]) : 0.f;
#pragma unroll
        for (int j = 0; j < MAXJ; j++) {
            const int cv = g.lane + g.lanes * j;
            if (j < g.nj && cv < g.cvecs) {
                const long long e = base + p * C + (long long)cv * VEC;
                Vec16<T> av, gv, yv, xv, o1, o2;
                if (ggx) av = ld16_stream(ggx + e);
                gv = ld16_stream(gy + e);
                yv = ld16_stream(yref + e);
                if (need_x) xv = ld16_stream(x + e);
                const long long nc = (long long)n * C + cv * VEC;
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    const float m = act_slope<ACT>((float)to_acc<T>(yv.v[k]), alpha, gain, clampv);
                    const float a = ggx ? (float)to_acc<T>(av.v[k]) : 0.f;
                    const float gdv = ggd ? (float)to_acc<T>(from_acc<T>(ggd[nc + k])) : 0.f;
                    const float g1 = (float)to_acc<T>(from_acc<T>((float)to_acc<T>(gv.v[k]) * m));   // materialised in T like the first-order pass
                    float t = nzv;
                    if (ggx) t += d ? a * (float)to_acc<T>(from_acc<T>(d[nc + k])) : a;
                    if (need_x) t += gdv * (float)to_acc<T>(xv.v[k]);
                    if (ggs0) t += (float)to_acc<T>(from_acc<T>(ggs0[nc + k]));
                    o1.v[k] = from_acc<T>(m * t);
                    if (d_x) o2.v[k] = from_acc<T>(gdv * g1);
                    if (partial) acc[j][k] += a * g1;
                }
                if (d_gy) st16_stream(d_gy + e, o1);
                if (d_x) st16_stream(d_x + e, o2);
            }
        }
    }
    if (partial) write_partial<VEC>(g, acc, red, partial + ((long long)n * bands + band) * C, C);
}

// out[n][c] = sum over bands (in order) of partial[n][band][c]
__global__ void band_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int NC_n, int C, int bands) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)NC_n * C) return;
    const long long n = i / C;
    const int c = (int)(i - n * C);
    float sum = 0.f;
    for (int b = 0; b < bands; b++) sum += partial[(n * bands + b) * C + c];
    out[i] = sum;
}

int pick_bands(int N, long long P, int C, int vec) {
    const int cvecs = C / vec;
    const int lanes = cvecs < 32 ? cvecs : 32;
    const int rgroups = 256 / lanes;
    long long bands = ((long long)gt_num_sms() * 8 + N - 1) / N;        // about 8 CTAs per SM in total
    const long long maxb = (P + rgroups - 1) / rgroups;
    if (bands > maxb) bands = maxb;
    if (bands < 1) bands = 1;
    return (int)bands;
}

int ew_grid(long long total_vecs) {
    long long blocks = (total_vecs + 255) / 256;
    const long long cap = (long long)gt_num_sms() * 16;
    return (int)(blocks < cap ? blocks : cap);
}

template <class T>
bool shape_ok(int N, long long P, int C) {
    constexpr int VEC = Vec16<T>::N;
    return N > 0 && P > 0 && C > 0 && C % VEC == 0 && C / VEC <= 32 * MAXJ && (C / VEC >= 32 ? (C / VEC) % 32 == 0 : (256 % (C / VEC)) == 0);
}

}  // namespace

#define GT_MOD_DISPATCH(CALL_F32, CALL_F16)                                   \
    if (dtype == GT_F32) { CALL_F32; }                                        \
    else if (dtype == GT_F16) { CALL_F16; }                                   \
    else { gt_set_error("modulated ops: unsupported dtype code %d", dtype); return GT_ERR_ARG; }

extern "C" long long gt_mod_workspace(int N, long long P, int C, int dtype) {
    const int vec = dtype == GT_F16 ? 8 : 4;
    if (N <= 0 || P <= 0 || C <= 0 || C % vec) return 0;
    return 2ll * N * pick_bands(N, P, C, vec) * C;
}

extern "C" int gt_mod_scale_fwd(const void* x, const float* s, void* y, int dtype, int N, long long P, int C, void* stream) {
    GT_REQUIRE(x && s && y, "gt_mod_scale_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = dtype == GT_F16 ? 8 : 4;
    GT_REQUIRE(N > 0 && P > 0 && C > 0 && C % vec == 0, "gt_mod_scale_fwd: bad shape N=%d P=%lld C=%d", N, P, C);
#define GT_MSF_BULK(T_)                                                                                                                    \
    if (bulk_ok<T_>(N, P, C, x, y, nullptr, nullptr)) {                                                                                    \
        constexpr int SMEM = streamk::Smem<1, 6, 16384>::TOTAL;                                                                            \
        static bool configured = false;                                                                                                    \
        if (!configured) {                                                                                                                 \
            int rc = set_smem(mod_scale_fwd_bulk_kernel<T_>, SMEM, "gt_mod_scale_fwd(bulk)");                                              \
            if (rc != GT_OK) return rc;                                                                                                    \
            configured = true;                                                                                                             \
        }                                                                                                                                  \
        dim3 grid(bulk_bands(N, P, C, (int)sizeof(T_), 16384, 1 << 30), N);                                                                \
        mod_scale_fwd_bulk_kernel<T_><<<grid, streamk::NTHREADS, SMEM, st>>>((const T_*)x, s, (T_*)y, C, P);                                \
        GT_CUDA_LAUNCH_CHECK("gt_mod_scale_fwd(bulk)");                                                                                    \
        return GT_OK;                                                                                                                      \
    }
    if (dtype == GT_F16) { GT_MSF_BULK(__half) } else if (dtype == GT_F32) { GT_MSF_BULK(float) }
#undef GT_MSF_BULK
    const long long tv = (long long)N * P * (C / vec);
    GT_MOD_DISPATCH((mod_scale_fwd_kernel<float><<<ew_grid(tv), 256, 0, st>>>((const float*)x, s, (float*)y, C, P, tv)),
                    (mod_scale_fwd_kernel<__half><<<ew_grid(tv), 256, 0, st>>>((const __half*)x, s, (__half*)y, C, P, tv)));
    GT_CUDA_LAUNCH_CHECK("gt_mod_scale_fwd");
    return GT_OK;
}

extern "C" int gt_mod_scale_bwd(const void* gy, const void* x, const float* s, void* gx, float* gs, float* workspace, long long workspace_floats,
                                int dtype, int N, long long P, int C, void* stream) {
    GT_REQUIRE(gy && x && s && gx && gs && workspace, "gt_mod_scale_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = dtype == GT_F16 ? 8 : 4;
    GT_REQUIRE((dtype == GT_F16 && shape_ok<__half>(N, P, C)) || (dtype == GT_F32 && shape_ok<float>(N, P, C)), "gt_mod_scale_bwd: unsupported shape N=%d P=%lld C=%d",
               N, P, C);
    const int bands_alloc = pick_bands(N, P, C, vec);
#define GT_MSB_BULK(T_)                                                                                                                    \
    if (bulk_ok<T_>(N, P, C, gy, x, gx, nullptr)) {                                                                                        \
        constexpr int SMEM = streamk::Smem<2, 6, 8192>::TOTAL;                                                                             \
        static bool configured = false;                                                                                                    \
        if (!configured) {                                                                                                                 \
            int rc = set_smem(mod_scale_bwd_bulk_kernel<T_>, SMEM, "gt_mod_scale_bwd(bulk)");                                              \
            if (rc != GT_OK) return rc;                                                                                                    \
            configured = true;                                                                                                             \
        }                                                                                                                                  \
        const int bb = bulk_bands(N, P, C, (int)sizeof(T_), 8192, bands_alloc);                                                            \
        GT_REQUIRE((long long)N * bb * C <= workspace_floats, "gt_mod_scale_bwd: workspace too small");                                    \
        dim3 grid(bb, N);                                                                                                                  \
        mod_scale_bwd_bulk_kernel<T_><<<grid, streamk::NTHREADS, SMEM, st>>>((const T_*)gy, (const T_*)x, s, (T_*)gx, workspace, C, P);     \
        GT_CUDA_LAUNCH_CHECK("gt_mod_scale_bwd(bulk)");                                                                                    \
        band_reduce_kernel<<<(int)(((long long)N * C + 255) / 256), 256, 0, st>>>(workspace, gs, N, C, bb);                                 \
        GT_CUDA_LAUNCH_CHECK("gt_mod_scale_bwd(reduce)");                                                                                  \
        return GT_OK;                                                                                                                      \
    }
    if (dtype == GT_F16) { GT_MSB_BULK(__half) } else if (dtype == GT_F32) { GT_MSB_BULK(float) }
#undef GT_MSB_BULK
    const int bands = bands_alloc;
    GT_REQUIRE((long long)N * bands * C <= workspace_floats, "gt_mod_scale_bwd: workspace too small");
    dim3 grid(bands, N);
    GT_MOD_DISPATCH((mod_scale_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)gy, (const float*)x, s, (float*)gx, workspace, C, P)),
                    (mod_scale_bwd_kernel<__half><<<grid, 256, 0, st>>>((const __half*)gy, (const __half*)x, s, (__half*)gx, workspace, C, P)));
    GT_CUDA_LAUNCH_CHECK("gt_mod_scale_bwd");
    band_reduce_kernel<<<(int)(((long long)N * C + 255) / 256), 256, 0, st>>>(workspace, gs, N, C, bands);
    GT_CUDA_LAUNCH_CHECK("gt_mod_scale_bwd(reduce)");
    return GT_OK;
}

extern "C" int gt_demod_act_fwd(const void* x, const float* d, const void* noise, const void* b, void* y, int dtype, int act, float alpha, float gain,
                                float clamp, int N, long long P, int C, void* stream) {
    GT_REQUIRE(x && y, "gt_demod_act_fwd: null pointer");
    GT_REQUIRE(act == A_LINEAR || act == A_LRELU, "gt_demod_act_fwd: only linear (1) and lrelu (3) are supported; got %d", act);
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = dtype == GT_F16 ? 8 : 4;
    GT_REQUIRE(N > 0 && P > 0 && C > 0 && C % vec == 0, "gt_demod_act_fwd: bad shape N=%d P=%lld C=%d", N, P, C);
#define GT_DAF_BULK(T_, ACT_)                                                                                                              \
    {                                                                                                                                      \
        constexpr int SMEM = streamk::Smem<1, 6, 16384>::TOTAL;                                                                            \
        static bool configured = false;                                                                                                    \
        if (!configured) {                                                                                                                 \
            int rc = set_smem(demod_act_fwd_bulk_kernel<T_, ACT_>, SMEM, "gt_demod_act_fwd(bulk)");                                        \
            if (rc != GT_OK) return rc;                                                                                                    \
            configured = true;                                                                                                             \
        }                                                                                                                                  \
        dim3 grid(bulk_bands(N, P, C, (int)sizeof(T_), 16384, 1 << 30), N);                                                                \
        demod_act_fwd_bulk_kernel<T_, ACT_><<<grid, streamk::NTHREADS, SMEM, st>>>((const T_*)x, d, (const T_*)noise, (const T_*)b, (T_*)y, C, ilog2(C), P, alpha, gain, clamp); \
        GT_CUDA_LAUNCH_CHECK("gt_demod_act_fwd(bulk)");                                                                                    \
        return GT_OK;                                                                                                                      \
    }
    // the per-pixel noise rides along as the side stream: 16384 / C bytes per chunk (16 .. 512), 16-byte aligned per sample
    const bool noise_ok = noise == nullptr || (C >= 32 && C <= 1024 && (((uintptr_t)noise) & 15) == 0 && (P * (dtype == GT_F16 ? 2 : 4)) % 16 == 0);
    if (!noise_ok) {
    } else if (dtype == GT_F16 && bulk_ok<__half>(N, P, C, x, y, b, nullptr)) {
        if (act == A_LRELU) GT_DAF_BULK(__half, A_LRELU) else GT_DAF_BULK(__half, A_LINEAR)
    } else if (dtype == GT_F32 && bulk_ok<float>(N, P, C, x, y, b, nullptr)) {
        if (act == A_LRELU) GT_DAF_BULK(float, A_LRELU) else GT_DAF_BULK(float, A_LINEAR)
    }
#undef GT_DAF_BULK
    const long long tv = (long long)N * P * (C / vec);
    const int grid = ew_grid(tv);
    if (act == A_LRELU) {
        GT_MOD_DISPATCH((demod_act_fwd_kernel<float, A_LRELU><<<grid, 256, 0, st>>>((const float*)x, d, (const float*)noise, (const float*)b, (float*)y, C, P, tv, alpha, gain, clamp)),
                        (demod_act_fwd_kernel<__half, A_LRELU><<<grid, 256, 0, st>>>((const __half*)x, d, (const __half*)noise, (const __half*)b, (__half*)y, C, P, tv, alpha, gain, clamp)));
    } else {
        GT_MOD_DISPATCH((demod_act_fwd_kernel<float, A_LINEAR><<<grid, 256, 0, st>>>((const float*)x, d, (const float*)noise, (const float*)b, (float*)y, C, P, tv, alpha, gain, clamp)),
                        (demod_act_fwd_kernel<__half, A_LINEAR><<<grid, 256, 0, st>>>((const __half*)x, d, (const __half*)noise, (const __half*)b, (__half*)y, C, P, tv, alpha, gain, clamp)));
    }
    GT_CUDA_LAUNCH_CHECK("gt_demod_act_fwd");
    return GT_OK;
}

extern "C" int gt_demod_act_bwd(const void* gy, const void* yref, const void* x, const float* d, void* gx, void* gnoise, float* gd, float* s0,
                                float* workspace, long long workspace_floats, int dtype, int act, float alpha, float gain, float clamp, int N,
                                long long P, int C, void* stream) {
    GT_REQUIRE(gy && yref && gx && s0 && workspace, "gt_demod_act_bwd: null pointer");
    GT_REQUIRE((d == nullptr) == (gd == nullptr) && (d == nullptr || x != nullptr), "gt_demod_act_bwd: d, gd and x go together");
    GT_REQUIRE(act == A_LINEAR || act == A_LRELU, "gt_demod_act_bwd: only linear (1) and lrelu (3) are supported; got %d", act);
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = dtype == GT_F16 ? 8 : 4;
    GT_REQUIRE((dtype == GT_F16 && shape_ok<__half>(N, P, C)) || (dtype == GT_F32 && shape_ok<float>(N, P, C)), "gt_demod_act_bwd: unsupported shape N=%d P=%lld C=%d",
               N, P, C);
    const int bands_alloc = pick_bands(N, P, C, vec);
#define GT_DAB_BULK(T_, ACT_, HASD_)                                                                                                       \
    {                                                                                                                                      \
        constexpr int SMEM = HASD_ ? streamk::Smem<3, 4, 8192>::TOTAL : streamk::Smem<2, 6, 8192>::TOTAL;                                   \
        static bool configured = false;                                                                                                    \
        if (!configured) {                                                                                                                 \
            int rc = set_smem(demod_act_bwd_bulk_kernel<T_, ACT_, HASD_>, SMEM, "gt_demod_act_bwd(bulk)");                                  \
            if (rc != GT_OK) return rc;                                                                                                    \
            configured = true;                                                                                                             \
        }                                                                                                                                  \
        const int bb = bulk_bands(N, P, C, (int)sizeof(T_), 8192, bands_alloc);                                                            \
        const long long perb = (long long)N * bb * C;                                                                                      \
        GT_REQUIRE(2 * perb <= workspace_floats, "gt_demod_act_bwd: workspace too small");                                                 \
        float* q1 = HASD_ ? workspace : nullptr;                                                                                           \
        float* q0 = workspace + perb;                                                                                                      \
        dim3 grid(bb, N);                                                                                                                  \
        demod_act_bwd_bulk_kernel<T_, ACT_, HASD_><<<grid, streamk::NTHREADS, SMEM, st>>>((const T_*)gy, (const T_*)yref, (const T_*)x, d, (T_*)gx, (T_*)gnoise, q1, q0, C, ilog2(C), P, alpha, gain, clamp); \
        GT_CUDA_LAUNCH_CHECK("gt_demod_act_bwd(bulk)");                                                                                    \
        const int rg = (int)(((long long)N * C + 255) / 256);                                                                              \
        if (HASD_) {                                                                                                                       \
            band_reduce_kernel<<<rg, 256, 0, st>>>(q1, gd, N, C, bb);                                                                       \
            GT_CUDA_LAUNCH_CHECK("gt_demod_act_bwd(reduce gd)");                                                                           \
        }                                                                                                                                  \
        band_reduce_kernel<<<rg, 256, 0, st>>>(q0, s0, N, C, bb);                                                                           \
        GT_CUDA_LAUNCH_CHECK("gt_demod_act_bwd(reduce s0)");                                                                               \
        return GT_OK;                                                                                                                      \
    }
#define GT_DAB_BULK_T(T_)                                                                                                                  \
    if (bulk_ok<T_>(N, P, C, gy, yref, x, gx) && (gnoise == nullptr || C / (int)Vec16<T_>::N <= 32)) {                                      \
        if (act == A_LRELU) {                                                                                                              \
            if (d) GT_DAB_BULK(T_, A_LRELU, true) else GT_DAB_BULK(T_, A_LRELU, false)                                                     \
        } else {                                                                                                                           \
            if (d) GT_DAB_BULK(T_, A_LINEAR, true) else GT_DAB_BULK(T_, A_LINEAR, false)                                                   \
        }                                                                                                                                  \
    }
    if (dtype == GT_F16) { GT_DAB_BULK_T(__half) } else if (dtype == GT_F32) { GT_DAB_BULK_T(float) }
#undef GT_DAB_BULK_T
#undef GT_DAB_BULK
    const int bands = bands_alloc;
    const long long per = (long long)N * bands * C;
    GT_REQUIRE(2 * per <= workspace_floats, "gt_demod_act_bwd: workspace too small");
    float* p1 = d ? workspace : nullptr;
    float* p0 = workspace + per;
    dim3 grid(bands, N);
    if (act == A_LRELU) {
        GT_MOD_DISPATCH((demod_act_bwd_kernel<float, A_LRELU><<<grid, 256, 0, st>>>((const float*)gy, (const float*)yref, (const float*)x, d, (float*)gx, (float*)gnoise, p1, p0, C, P, alpha, gain, clamp)),
                        (demod_act_bwd_kernel<__half, A_LRELU><<<grid, 256, 0, st>>>((const __half*)gy, (const __half*)yref, (const __half*)x, d, (__half*)gx, (__half*)gnoise, p1, p0, C, P, alpha, gain, clamp)));
    } else {
        GT_MOD_DISPATCH((demod_act_bwd_kernel<float, A_LINEAR><<<grid, 256, 0, st>>>((const float*)gy, (const float*)yref, (const float*)x, d, (float*)gx, (float*)gnoise, p1, p0, C, P, alpha, gain, clamp)),
                        (demod_act_bwd_kernel<__half, A_LINEAR><<<grid, 256, 0, st>>>((const __half*)gy, (const __half*)yref, (const __half*)x, d, (__half*)gx, (__half*)gnoise, p1, p0, C, P, alpha, gain, clamp)));
    }
    GT_CUDA_LAUNCH_CHECK("gt_demod_act_bwd");
    const int rgrid = (int)(((long long)N * C + 255) / 256);
    if (d) {
        band_reduce_kernel<<<rgrid, 256, 0, st>>>(p1, gd, N, C, bands);
        GT_CUDA_LAUNCH_CHECK("gt_demod_act_bwd(reduce gd)");
    }
    band_reduce_kernel<<<rgrid, 256, 0, st>>>(p0, s0, N, C, bands);
    GT_CUDA_LAUNCH_CHECK("gt_demod_act_bwd(reduce s0)");
    return GT_OK;
}

extern "C" int gt_mod_scale_bwd2(const void* ggx, const float* ggs, const void* gy, const void* x, const float* s, void* d_gy, void* d_x, float* d_s,
                                 float* workspace, long long workspace_floats, int dtype, int N, long long P, int C, void* stream) {
    GT_REQUIRE(gy && x && s && workspace, "gt_mod_scale_bwd2: null pointer");
    GT_REQUIRE(d_s == nullptr || ggx != nullptr, "gt_mod_scale_bwd2: d_s needs ggx");
    GT_REQUIRE(d_x == nullptr || ggs != nullptr, "gt_mod_scale_bwd2: d_x needs ggs");
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = dtype == GT_F16 ? 8 : 4;
    GT_REQUIRE((dtype == GT_F16 && shape_ok<__half>(N, P, C)) || (dtype == GT_F32 && shape_ok<float>(N, P, C)), "gt_mod_scale_bwd2: unsupported shape N=%d P=%lld C=%d",
               N, P, C);
    const int bands = pick_bands(N, P, C, vec);
    GT_REQUIRE((long long)N * bands * C <= workspace_floats, "gt_mod_scale_bwd2: workspace too small");
    dim3 grid(bands, N);
    float* part = d_s ? workspace : nullptr;
    GT_MOD_DISPATCH((mod_scale_bwd2_kernel<float><<<grid, 256, 0, st>>>((const float*)ggx, ggs, (const float*)gy, (const float*)x, s, (float*)d_gy, (float*)d_x, part, C, P)),
                    (mod_scale_bwd2_kernel<__half><<<grid, 256, 0, st>>>((const __half*)ggx, ggs, (const __half*)gy, (const __half*)x, s, (__half*)d_gy, (__half*)d_x, part, C, P)));
    GT_CUDA_LAUNCH_CHECK("gt_mod_scale_bwd2");
    if (d_s) {
        band_reduce_kernel<<<(int)(((long long)N * C + 255) / 256), 256, 0, st>>>(workspace, d_s, N, C, bands);
        GT_CUDA_LAUNCH_CHECK("gt_mod_scale_bwd2(reduce)");
    }
    return GT_OK;
}

extern "C" int gt_demod_act_bwd2(const void* ggx, const float* ggd, const void* ggnz, const float* ggs0, const void* gy, const void* yref, const void* x,
                                 const float* d, void* d_gy, void* d_x, float* d_d, float* workspace, long long workspace_floats, int dtype, int act,
                                 float alpha, float gain, float clamp, int N, long long P, int C, void* stream) {
    GT_REQUIRE(gy && yref && workspace, "gt_demod_act_bwd2: null pointer");
    GT_REQUIRE(act == A_LINEAR || act == A_LRELU, "gt_demod_act_bwd2: only linear (1) and lrelu (3) are supported; got %d", act);
    GT_REQUIRE(d_d == nullptr || (ggx != nullptr && d != nullptr), "gt_demod_act_bwd2: d_d needs ggx and d");
    GT_REQUIRE(d_x == nullptr || ggd != nullptr, "gt_demod_act_bwd2: d_x needs ggd");
    GT_REQUIRE(ggd == nullptr || x != nullptr || d_gy == nullptr, "gt_demod_act_bwd2: ggd needs x");
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = dtype == GT_F16 ? 8 : 4;
    GT_REQUIRE((dtype == GT_F16 && shape_ok<__half>(N, P, C)) || (dtype == GT_F32 && shape_ok<float>(N, P, C)), "gt_demod_act_bwd2: unsupported shape N=%d P=%lld C=%d",
               N, P, C);
    const int bands = pick_bands(N, P, C, vec);
    GT_REQUIRE((long long)N * bands * C <= workspace_floats, "gt_demod_act_bwd2: workspace too small");
    dim3 grid(bands, N);
    float* part = d_d ? workspace : nullptr;
    if (act == A_LRELU) {
        GT_MOD_DISPATCH((demod_act_bwd2_kernel<float, A_LRELU><<<grid, 256, 0, st>>>((const float*)ggx, ggd, (const float*)ggnz, ggs0, (const float*)gy, (const float*)yref, (const float*)x, d, (float*)d_gy, (float*)d_x, part, C, P, alpha, gain, clamp)),
                        (demod_act_bwd2_kernel<__half, A_LRELU><<<grid, 256, 0, st>>>((const __half*)ggx, ggd, (const __half*)ggnz, ggs0, (const __half*)gy, (const __half*)yref, (const __half*)x, d, (__half*)d_gy, (__half*)d_x, part, C, P, alpha, gain, clamp)));
    } else {
        GT_MOD_DISPATCH((demod_act_bwd2_kernel<float, A_LINEAR><<<grid, 256, 0, st>>>((const float*)ggx, ggd, (const float*)ggnz, ggs0, (const float*)gy, (const float*)yref, (const float*)x, d, (float*)d_gy, (float*)d_x, part, C, P, alpha, gain, clamp)),
                        (demod_act_bwd2_kernel<__half, A_LINEAR><<<grid, 256, 0, st>>>((const __half*)ggx, ggd, (const __half*)ggnz, ggs0, (const __half*)gy, (const __half*)yref, (const __half*)x, d, (__half*)d_gy, (__half*)d_x, part, C, P, alpha, gain, clamp)));
    }
    GT_CUDA_LAUNCH_CHECK("gt_demod_act_bwd2");
    if (d_d) {
        band_reduce_kernel<<<(int)(((long long)N * C + 255) / 256), 256, 0, st>>>(workspace, d_d, N, C, bands);
        GT_CUDA_LAUNCH_CHECK("gt_demod_act_bwd2(reduce)");
    }
    return GT_OK;
}
