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
    const int cvecs = C / VEC;                                   // divides 256 (launcher), so a thread's channels are fixed
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
        const long long pix = i / cvecs;
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
            const long long pix = i / cvecs;
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
            if (live && cv == 0) dx[i / cvecs] = from_acc<T>(s);
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
