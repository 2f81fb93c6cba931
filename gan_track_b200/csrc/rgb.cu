// rgb.cu -- the single-channel image <-> feature-map layers at the top resolution.
//
// FromRGB of the discriminator (S3/training/networks_stylegan2.py:586, 617-621; S3 = /root/reference/src/models/stylegan3) is a
// Conv2dLayer(img_channels -> C, kernel 1, lrelu): with ONE image channel it is an outer product per pixel followed by
// bias_act.  Routed through a library convolution it costs a GEMM-shaped kernel, a separate bias_act pass over the
// [N,C,H,W] result and -- because the library returns NCHW for a 1-channel input -- a full layout-conversion copy before the
// first tcgen05 convolution (1.4 ms / step of strided copies in profiles/r01b_step_profile.txt).  Here:
//     gt_fromrgb1_fwd   y[n,p,c] = clamp(act(round_T(x[n,p] * w[c]) + b[c]) * gain)           written channels-last, once
//     gt_fromrgb1_bwd   g1 = dy * gain * act'(y) * [|y| < clamp];  dw[c] = sum g1 x;  db[c] = sum g1;  dx[n,p] = sum_c g1 w[c]
// one pass each.  (Second-order use -- R1 differentiates this backward -- takes the tensor-op formulation in the binding.)
#include "gt_common.cuh"
#include "hot_act.cuh"

namespace {

template <class T, int ACT>
__global__ void __launch_bounds__(256) fromrgb1_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b, T* __restrict__ y,
                                                           long long NP, int C, float alpha, float gain, float clampv) {
    typedef hot::Lanes<T> L;
    constexpr int VEC = Vec16<T>::N;
    const int cvecs = C / VEC;                                   // a power of two dividing 256 (launcher), so a thread's channels are fixed
    const int lcv = 31 - __clz(cvecs);
    const int cv = threadIdx.x % cvecs;
    const hot::Params hp = hot::make_params(alpha, gain, clampv);
    const bool clamp_on = clampv >= 0.f;
    float2 wv[L::NP], bv[L::NP];
    {
        const Vec16<T> wr = ld16(w + cv * VEC);
#pragma unroll
        for (int i = 0; i < L::NP; i++) wv[i] = L::get(wr, i), bv[i] = make_float2(0.f, 0.f);
        if (b) {
            const Vec16<T> br = ld16(b + cv * VEC);
#pragma unroll
            for (int i = 0; i < L::NP; i++) bv[i] = L::get(br, i);
        }
    }
    const long long total = NP * cvecs;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long pix = i >> lcv;
        const float xv = (float)to_acc<T>(x[pix]);
        Vec16<T> o;
#pragma unroll
        for (int k = 0; k < L::NP; k++) {
            float2 u = __fmul2_rn(make_float2(xv, xv), wv[k]);
            L::set(o, k, u);                                     // the convolution output is materialised in T ...
            u = __fadd2_rn(L::get(o, k), bv[k]);                  // ... before bias_act reads it back
            L::set(o, k, clamp_on ? hot::fwd<ACT, true>(u, hp) : hot::fwd<ACT, false>(u, hp));
        }
        st16_stream(y + i * VEC, o);
    }
}

template <class T, int ACT>
__global__ void __launch_bounds__(256) fromrgb1_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                                                           const T* __restrict__ w, T* __restrict__ dx, float* __restrict__ partial, long long NP, int C,
                                                           float alpha, float gain, float clampv) {
    typedef hot::Lanes<T> L;
    constexpr int VEC = Vec16<T>::N;
    __shared__ float red[256 * VEC];
    const int cvecs = C / VEC;                                   // a power of two <= 32 (launcher): one pixel = adjacent lanes of a warp
    const int lcv = 31 - __clz(cvecs);
    const int cv = threadIdx.x % cvecs;
    const hot::Params hp = hot::make_params(alpha, gain, clampv);
    const bool clamp_on = clampv >= 0.f;
    float2 wv[L::NP];
    {
        const Vec16<T> wr = ld16(w + cv * VEC);
#pragma unroll
        for (int i = 0; i < L::NP; i++) wv[i] = L::get(wr, i);
    }
    float aw[VEC], ab[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) aw[k] = ab[k] = 0.f;
    const long long total = NP * cvecs;
    const long long iters = (total + (long long)gridDim.x * 256 - 1) / ((long long)gridDim.x * 256);
    for (long long it = 0; it < iters; it++) {
        const long long i = ((long long)it * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        const bool live = i < total;
        float s = 0.f;
        if (live) {
            const long long pix = (i >> lcv);
            const float xv = (float)to_acc<T>(x[pix]);
            const Vec16<T> g = ld16_stream(dy + i * VEC), yv = ld16_stream(y + i * VEC);
#pragma unroll
            for (int k = 0; k < L::NP; k++) {
                float2 g1 = clamp_on ? hot::bwd<ACT, true>(L::get(g, k), L::get(yv, k), hp) : hot::bwd<ACT, false>(L::get(g, k), L::get(yv, k), hp);
                Vec16<T> tmp;
                L::set(tmp, 0, g1);
                g1 = L::get(tmp, 0);                             // materialised in T like bias_act's gradient output
                aw[2 * k] += g1.x * xv;
                aw[2 * k + 1] += g1.y * xv;
                ab[2 * k] += g1.x;
                ab[2 * k + 1] += g1.y;
                s += g1.x * wv[k].x + g1.y * wv[k].y;
            }
        }
        if (dx) {
            for (int off = cvecs >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (live && cv == 0) dx[(i >> lcv)] = from_acc<T>(s);
        }
    }
    // per-channel sums of this CTA in a fixed order: dw then db -> partial[block][2][C]
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < VEC; k++) red[threadIdx.x * VEC + k] = pass == 0 ? aw[k] : ab[k];
        __syncthreads();
        if ((int)threadIdx.x < cvecs) {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                float sum = 0.f;
                for (int t = threadIdx.x; t < 256; t += cvecs) sum += red[t * VEC + k];
                partial[((long long)blockIdx.x * 2 + pass) * C + threadIdx.x * VEC + k] = sum;
            }
        }
    }
}

// out[j][c] = sum over blocks (in order) of partial[block][j][c], j in {0: dw, 1: db}
__global__ void fromrgb1_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db, int C, int blocks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * C) return;
    const int j = i / C, c = i - j * C;
    float s = 0.f;
    for (int b = 0; b < blocks; b++) s += partial[((long long)b * 2 + j) * C + c];
    (j == 0 ? dw : db)[c] = s;
}

template <class T>
bool shape_ok(int C) {
    constexpr int VEC = Vec16<T>::N;
    const int cv = C / VEC;
    return C % VEC == 0 && cv >= 1 && cv <= 32 && (cv & (cv - 1)) == 0;
}

int bwd_grid() { return gt_num_sms() * 4; }

}  // namespace

extern "C" long long gt_fromrgb1_bwd_workspace(int C) { return (long long)bwd_grid() * 2 * C; }

extern "C" int gt_fromrgb1_fwd(const void* x, const void* w, const void* b, void* y, int dtype, int act, float alpha, float gain, float clamp, long long NP,
                               int C, void* stream) {
    GT_REQUIRE(x && w && y, "gt_fromrgb1_fwd: null pointer");
    GT_REQUIRE(act == 1 || act == 3, "gt_fromrgb1_fwd: act must be linear (1) or lrelu (3); got %d", act);
    GT_REQUIRE(NP > 0 && ((dtype == GT_F16 && shape_ok<__half>(C)) || (dtype == GT_F32 && shape_ok<float>(C))), "gt_fromrgb1_fwd: unsupported shape C=%d", C);
    GT_REQUIRE(((((uintptr_t)w) | ((uintptr_t)b) | ((uintptr_t)y)) & 15) == 0, "gt_fromrgb1_fwd: w, b, y must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = dtype == GT_F16 ? 8 : 4;
    long long blocks = (NP * (C / vec) + 255) / 256;
    if (blocks > (long long)gt_num_sms() * 16) blocks = (long long)gt_num_sms() * 16;
#define LAUNCH(T_, ACT_) fromrgb1_fwd_kernel<T_, ACT_><<<(int)blocks, 256, 0, st>>>((const T_*)x, (const T_*)w, (const T_*)b, (T_*)y, NP, C, alpha, gain, clamp)
    if (dtype == GT_F16) { if (act == 3) LAUNCH(__half, 3); else LAUNCH(__half, 1); }
    else { if (act == 3) LAUNCH(float, 3); else LAUNCH(float, 1); }
#undef LAUNCH
    GT_CUDA_LAUNCH_CHECK("gt_fromrgb1_fwd");
    return GT_OK;
}

extern "C" int gt_fromrgb1_bwd(const void* dy, const void* y, const void* x, const void* w, void* dx, float* dw, float* db, float* workspace,
                               long long workspace_floats, int dtype, int act, float alpha, float gain, float clamp, long long NP, int C, void* stream) {
    GT_REQUIRE(dy && y && x && w && dw && db && workspace, "gt_fromrgb1_bwd: null pointer");
    GT_REQUIRE(act == 1 || act == 3, "gt_fromrgb1_bwd: act must be linear (1) or lrelu (3); got %d", act);
    GT_REQUIRE(NP > 0 && ((dtype == GT_F16 && shape_ok<__half>(C)) || (dtype == GT_F32 && shape_ok<float>(C))), "gt_fromrgb1_bwd: unsupported shape C=%d", C);
    const int grid = bwd_grid();
    GT_REQUIRE(workspace_floats >= (long long)grid * 2 * C, "gt_fromrgb1_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(T_, ACT_) fromrgb1_bwd_kernel<T_, ACT_><<<grid, 256, 0, st>>>((const T_*)dy, (const T_*)y, (const T_*)x, (const T_*)w, (T_*)dx, workspace, NP, C, alpha, gain, clamp)
    if (dtype == GT_F16) { if (act == 3) LAUNCH(__half, 3); else LAUNCH(__half, 1); }
    else { if (act == 3) LAUNCH(float, 3); else LAUNCH(float, 1); }
#undef LAUNCH
    GT_CUDA_LAUNCH_CHECK("gt_fromrgb1_bwd");
    fromrgb1_reduce_kernel<<<(2 * C + 127) / 128, 128, 0, st>>>(workspace, dw, db, C, grid);
    GT_CUDA_LAUNCH_CHECK("gt_fromrgb1_bwd(reduce)");
    return GT_OK;
}

// =====================================================================================================================
// ToRGB with ONE image channel (S3/training/networks_stylegan2.py:338-358): a modulated 1x1 convolution without
// demodulation followed by a linear bias_act with clamp,
//     y[n,p] = clamp(round_T(sum_c round_T(x[n,p,c] * s[n,c]) * w[c]) + b)
// The op-by-op route writes the modulated activations (a full read + write pass), runs a library GEMM-shaped kernel with one
// output column and a bias_act over the image; here x is read once.  Backward (first order) in one pass:
//     g = dy (linear, gain 1; `linear` saves no output so the clamp does not mask, OPS/bias_act.py:151-154)
//     dxs = round_T(g w[c]);  dx = round_T(dxs s[n,c]);  ds[n,c] = sum_p dxs x;  dw[c] = sum_{n,p} g round_T(x s);  db = sum g
// x, dx: [N,P,C] channels-last; s, ds: [N,C] fp32; w: [C] T; y, dy: [N,P] T.
// =====================================================================================================================
namespace {

template <class T> __device__ __forceinline__ float2 rnd_io(float2 v);
template <> __device__ __forceinline__ float2 rnd_io<__half>(float2 v) { return __half22float2(__float22half2_rn(v)); }
template <> __device__ __forceinline__ float2 rnd_io<float>(float2 v) { return v; }
template <class T> __device__ __forceinline__ float rnd_io1(float v) { return (float)to_acc<T>(from_acc<T>(v)); }

template <class T>
__global__ void __launch_bounds__(256) torgb1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ s, const T* __restrict__ w, const T* __restrict__ b,
                                                         T* __restrict__ y, long long P, int C, float clampv) {
    typedef hot::Lanes<T> L;
    constexpr int VEC = Vec16<T>::N;
    const int cvecs = C / VEC;
    const int lcv = 31 - __clz(cvecs);
    const int cv = threadIdx.x % cvecs, n = blockIdx.y;
    float2 ws[L::NP], sv[L::NP];
    {
        const Vec16<T> wr = ld16(w + cv * VEC);
        const float* sp = s + (long long)n * C + cv * VEC;
#pragma unroll
        for (int i = 0; i < L::NP; i++) ws[i] = L::get(wr, i), sv[i] = rnd_io<T>(make_float2(sp[2 * i], sp[2 * i + 1]));
    }
    const float bv = b ? (float)to_acc<T>(b[0]) : 0.f;
    const long long total = P * cvecs;
    const T* xn = x + (long long)n * P * C;
    const long long iters = (total + (long long)gridDim.x * 256 - 1) / ((long long)gridDim.x * 256);
    for (long long it = 0; it < iters; it++) {
        const long long i = ((long long)it * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        const bool live = i < total;
        float dot = 0.f;
        if (live) {
            const Vec16<T> xv = ld16_stream(xn + i * VEC);
#pragma unroll
            for (int k = 0; k < L::NP; k++) {
                const float2 xs = rnd_io<T>(__fmul2_rn(L::get(xv, k), sv[k]));
                dot += xs.x * ws[k].x + xs.y * ws[k].y;
            }
        }
        for (int off = cvecs >> 1; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
        if (live && cv == 0) {
            float v = rnd_io1<T>(dot) + bv;
            if (clampv >= 0.f) v = hot::min_nan(hot::max_nan(v, -clampv), clampv);
            y[(long long)n * P + (i >> lcv)] = from_acc<T>(v);
        }
    }
}

// grid = (bands, N); partial[n][band][{ds, dw}][C] and partial_b[n][band]
template <class T>
__global__ void __launch_bounds__(256) torgb1_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ s, const T* __restrict__ w,
                                                         T* __restrict__ dx, float* __restrict__ partial, float* __restrict__ partial_b, long long P, int C) {
    typedef hot::Lanes<T> L;
    constexpr int VEC = Vec16<T>::N;
    __shared__ float red[256 * VEC];
    __shared__ float redb[256];
    const int cvecs = C / VEC;
    const int lcv = 31 - __clz(cvecs);
    const int cv = threadIdx.x % cvecs, n = blockIdx.y;
    float2 ws[L::NP], sv[L::NP];
    {
        const Vec16<T> wr = ld16(w + cv * VEC);
        const float* sp = s + (long long)n * C + cv * VEC;
#pragma unroll
        for (int i = 0; i < L::NP; i++) ws[i] = L::get(wr, i), sv[i] = rnd_io<T>(make_float2(sp[2 * i], sp[2 * i + 1]));
    }
    float as[VEC], aw[VEC], ab = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; k++) as[k] = aw[k] = 0.f;
    const long long total = P * cvecs;
    const T* xn = x + (long long)n * P * C;
    T* dxn = dx + (long long)n * P * C;
    const T* dyn = dy + (long long)n * P;
    auto body = [&](long long i, float g, const Vec16<T>& xv) {
        Vec16<T> o;
#pragma unroll
        for (int k = 0; k < L::NP; k++) {
            const float2 xf = L::get(xv, k);
            const float2 dxs = rnd_io<T>(__fmul2_rn(make_float2(g, g), ws[k]));
            const float2 xs = rnd_io<T>(__fmul2_rn(xf, sv[k]));
            L::set(o, k, __fmul2_rn(dxs, sv[k]));
            as[2 * k] += dxs.x * xf.x;
            as[2 * k + 1] += dxs.y * xf.y;
            aw[2 * k] += g * xs.x;
            aw[2 * k + 1] += g * xs.y;
        }
        if (cv == 0) ab += g;
        st16_stream(dxn + i * VEC, o);
    };
    // two vectors per iteration: twice the loads in flight per thread (the pass is latency-bound at 8 warps per CTA otherwise)
    const long long step = (long long)gridDim.x * 256;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    for (; i + step < total; i += 2 * step) {
        const float g0 = (float)to_acc<T>(dyn[(i >> lcv)]), g1 = (float)to_acc<T>(dyn[((i + step) >> lcv)]);
        const Vec16<T> x0 = ld16_stream(xn + i * VEC), x1 = ld16_stream(xn + (i + step) * VEC);
        body(i, g0, x0);
        body(i + step, g1, x1);
    }
    if (i < total) body(i, (float)to_acc<T>(dyn[(i >> lcv)]), ld16_stream(xn + i * VEC));
    float* out = partial + ((long long)n * gridDim.x + blockIdx.x) * 2 * C;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < VEC; k++) red[threadIdx.x * VEC + k] = pass == 0 ? as[k] : aw[k];
        __syncthreads();
        if ((int)threadIdx.x < cvecs) {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                float sum = 0.f;
                for (int t = threadIdx.x; t < 256; t += cvecs) sum += red[t * VEC + k];
                out[pass * C + threadIdx.x * VEC + k] = sum;
            }
        }
    }
    redb[threadIdx.x] = ab;
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
        for (int t = 0; t < 256; t++) sum += redb[t];
        partial_b[(long long)n * gridDim.x + blockIdx.x] = sum;
    }
}

// ds[n][c] = sum_band partial[n][band][0][c];  dw[c] = sum_n sum_band partial[n][band][1][c];  db = sum partial_b   (fixed order)
// one warp per output value (ds[n,c]: `bands` terms; dw[c], db: N * bands terms), lanes stride over the terms, fixed order
__global__ void __launch_bounds__(256) torgb1_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_b, float* __restrict__ ds,
                                                            float* __restrict__ dw, float* __restrict__ db, int N, int C, int bands) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int nds = N * C;
    float sum = 0.f;
    if (o < nds) {
        const int n = o / C, c = o - n * C;
        for (int b = lane; b < bands; b += 32) sum += partial[((long long)n * bands + b) * 2 * C + c];
    } else if (o < nds + C) {
        const int c = o - nds;
        for (int k = lane; k < N * bands; k += 32) sum += partial[(long long)k * 2 * C + C + c];
    } else if (o == nds + C) {
        for (int k = lane; k < N * bands; k += 32) sum += partial_b[k];
    } else {
        return;
    }
    sum = warp_sum(sum);
    if (lane == 0) {
        if (o < nds) ds[o] = sum;
        else if (o < nds + C) dw[o - nds] = sum;
        else db[0] = sum;
    }
}

int torgb_bands(int N) {
    int b = (gt_num_sms() * 4 + N - 1) / N;
    return b < 1 ? 1 : b;
}

}  // namespace

extern "C" long long gt_torgb1_bwd_workspace(int N, int C) { return (long long)N * torgb_bands(N) * (2 * C + 1); }

extern "C" int gt_torgb1_fwd(const void* x, const float* s, const void* w, const void* b, void* y, int dtype, float clamp, int N, long long P, int C,
                             void* stream) {
    GT_REQUIRE(x && s && w && y, "gt_torgb1_fwd: null pointer");
    GT_REQUIRE(N > 0 && N <= 65535 && P > 0 && ((dtype == GT_F16 && shape_ok<__half>(C)) || (dtype == GT_F32 && shape_ok<float>(C))),
               "gt_torgb1_fwd: unsupported shape N=%d C=%d", N, C);
    GT_REQUIRE(((((uintptr_t)x) | ((uintptr_t)w)) & 15) == 0, "gt_torgb1_fwd: x, w must be 16-byte aligned");
    const int vec = dtype == GT_F16 ? 8 : 4;
    long long blocks = (P * (C / vec) + 255) / 256;
    const long long cap = ((long long)gt_num_sms() * 8 + N - 1) / N;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, (unsigned)N);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GT_F16) torgb1_fwd_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, s, (const __half*)w, (const __half*)b, (__half*)y, P, C, clamp);
    else torgb1_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, s, (const float*)w, (const float*)b, (float*)y, P, C, clamp);
    GT_CUDA_LAUNCH_CHECK("gt_torgb1_fwd");
    return GT_OK;
}

extern "C" int gt_torgb1_bwd(const void* dy, const void* x, const float* s, const void* w, void* dx, float* ds, float* dw, float* db, float* workspace,
                             long long workspace_floats, int dtype, int N, long long P, int C, void* stream) {
    GT_REQUIRE(dy && x && s && w && dx && ds && dw && db && workspace, "gt_torgb1_bwd: null pointer");
    GT_REQUIRE(N > 0 && N <= 65535 && P > 0 && ((dtype == GT_F16 && shape_ok<__half>(C)) || (dtype == GT_F32 && shape_ok<float>(C))),
               "gt_torgb1_bwd: unsupported shape N=%d C=%d", N, C);
    const int bands = torgb_bands(N);
    GT_REQUIRE(workspace_floats >= (long long)N * bands * (2 * C + 1), "gt_torgb1_bwd: workspace too small");
    float* part = workspace;
    float* part_b = workspace + (long long)N * bands * 2 * C;
    dim3 grid((unsigned)bands, (unsigned)N);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GT_F16)
        torgb1_bwd_kernel<__half><<<grid, 256, 0, st>>>((const __half*)dy, (const __half*)x, s, (const __half*)w, (__half*)dx, part, part_b, P, C);
    else
        torgb1_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (const float*)x, s, (const float*)w, (float*)dx, part, part_b, P, C);
    GT_CUDA_LAUNCH_CHECK("gt_torgb1_bwd");
    const int total = N * C + C + 1;
    torgb1_reduce_kernel<<<(total + 7) / 8, 256, 0, st>>>(part, part_b, ds, dw, db, N, C, bands);
    GT_CUDA_LAUNCH_CHECK("gt_torgb1_bwd(reduce)");
    return GT_OK;
}
