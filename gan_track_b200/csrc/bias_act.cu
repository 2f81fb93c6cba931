// bias_act for sm_100a: y = clamp(act(x + b) * gain) and its first / second derivative passes.
//
// Replaces the reference plugin OPS/bias_act.{cpp,cu} (launcher bias_act.cpp:32-90, kernel bias_act.cu:23-146;
// OPS = /root/reference/src/models/stylegan3/torch_utils/ops).  Same contract: dense x, bias along one dim given
// by (index / step_b) % size_b, grad in {0,1,2}, fp16 I/O with fp32 math.  Different execution: the op is pure HBM
// streaming, so the hot activations (linear, lrelu) run 16-byte vector loads/stores with 4 independent vectors in
// flight per thread, L1 no-allocate hints, one bias fetch per vector instead of an integer div+mod per element,
// and a grid sized to a whole number of waves over the SMs.  All nine activations are available on a scalar
// path.  gt_bias_act_bwd additionally produces the per-channel bias gradient in the same pass (the reference
// re-reads dx with a separate torch reduction, OPS/bias_act.py:169-170).
#include "gt_common.cuh"
#include "stream_bulk.cuh"
#include "hot_act.cuh"
#include <math.h>

namespace {

struct BiasActParams {
    const void* x;      // differentiated input: x (grad 0), dy (grad 1), d_dx (grad 2)
    const void* b;      // [size_b] or null
    const void* xref;   // forward input, only for activations that save x (swish) when grad > 0
    const void* yref;   // forward output, only when grad > 0 and the activation / clamp needs it
    const void* dy;     // only grad == 2
    void* y;
    int grad;
    float alpha, gain, clamp;
    long long size_x;
    int size_b;
    long long step_b;
};

enum { A_LINEAR = 1, A_RELU, A_LRELU, A_TANH, A_SIGMOID, A_ELU, A_SELU, A_SOFTPLUS, A_SWISH };

// One element.  S is the math type.  `v` is the differentiated input with the bias already applied where it
// belongs (to v for grad 0, to xr for grad > 0).
template <int ACT, class S>
__device__ __forceinline__ S act_eval(S v, S xr, S yr, S dyv, int grad, S alpha, S gain, S clampv) {
    const S one = (S)1, two = (S)2;
    const S selu_s = (S)1.0507009873554804934193349852946, selu_a = (S)1.6732632423543772848170429916717;
    S yy = (gain != (S)0) ? yr / gain : (S)0;   // forward output before gain
    S r = (S)0;
    if (ACT == A_LINEAR) {
        r = (grad < 2) ? v : (S)0;
    } else if (ACT == A_RELU) {
        r = (grad == 0) ? (v > 0 ? v : (S)0) : (grad == 1) ? (yy > 0 ? v : (S)0) : (S)0;
    } else if (ACT == A_LRELU) {
        r = (grad == 0) ? (v > 0 ? v : v * alpha) : (grad == 1) ? (yy > 0 ? v : v * alpha) : (S)0;
    } else if (ACT == A_TANH) {
        r = (grad == 0) ? (S)tanh((double)v) : (grad == 1) ? v * (one - yy * yy) : v * (one - yy * yy) * (-two * yy);
    } else if (ACT == A_SIGMOID) {
        r = (grad == 0) ? one / (one + (S)exp((double)-v)) : (grad == 1) ? v * yy * (one - yy) : v * yy * (one - yy) * (one - two * yy);
    } else if (ACT == A_ELU) {
        r = (grad == 0) ? (v >= 0 ? v : (S)expm1((double)v)) : (grad == 1) ? (yy >= 0 ? v : v * (yy + one)) : (yy >= 0 ? (S)0 : v * (yy + one));
    } else if (ACT == A_SELU) {
        const S sa = selu_s * selu_a;
        r = (grad == 0) ? (v >= 0 ? selu_s * v : sa * (S)expm1((double)v)) : (grad == 1) ? (yy >= 0 ? v * selu_s : v * (yy + sa)) : (yy >= 0 ? (S)0 : v * (yy + sa));
    } else if (ACT == A_SOFTPLUS) {
        if (grad == 0) r = (v > (S)20) ? v : (S)log1p(exp((double)v));
        else { S c = (S)exp((double)-yy); r = (grad == 1) ? v * (one - c) : v * c * (one - c); }
    } else if (ACT == A_SWISH) {
        if (grad == 0) r = v / (one + (S)exp((double)-v));
        else {
            S c = (S)exp((double)xr), d = c + one;
            if (grad == 1) r = (xr > (S)40) ? v : v * c * (xr + d) / (d * d);
            else r = (xr > (S)40) ? (S)0 : v * c * (xr * (two - d) + two * d) / (d * d * d);
            yr = xr / (one + (S)exp((double)-xr)) * gain;
        }
    }
    r *= gain * dyv;
    if (clampv >= (S)0) {
        if (grad == 0) r = (r > clampv) ? clampv : (r < -clampv) ? -clampv : r;
        else r = (yr > -clampv && yr < clampv) ? r : (S)0;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Scalar path: any activation, any layout parameters.
// ---------------------------------------------------------------------------------------------------------------
template <class T, int ACT>
__global__ void __launch_bounds__(256) bias_act_scalar_kernel(BiasActParams p) {
    typedef typename Acc<T>::type S;
    const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* xr = (const T*)p.xref;
    const T* yr = (const T*)p.yref; const T* dy = (const T*)p.dy; T* y = (T*)p.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.size_x; i += (long long)gridDim.x * blockDim.x) {
        S v = to_acc<T>(x[i]);
        S bv = b ? to_acc<T>(b[(i / p.step_b) % p.size_b]) : (S)0;
        S xrv = xr ? to_acc<T>(xr[i]) : (S)0;
        S yrv = yr ? to_acc<T>(yr[i]) : (S)0;
        S dyv = dy ? to_acc<T>(dy[i]) : (S)1;
        if (p.grad == 0) v += bv; else xrv += bv;
        y[i] = from_acc<T>(act_eval<ACT, S>(v, xrv, yrv, dyv, p.grad, (S)p.alpha, (S)p.gain, (S)p.clamp));
    }
}


// One 16-byte vector of the hot path (linear / lrelu, grad 0 or 1) with the lean packed arithmetic of hot_act.cuh.
// v0: x (grad 0) or dy (grad 1), replaced by the result; yv: saved forward output (grad 1); bv: the vector's bias lanes
// as fp32 pairs (grad 0 only, may be zeros).
template <class T, int ACT>
__device__ __forceinline__ void hot_vector(Vec16<T>& v0, const Vec16<T>& yv, const float2* bv, int grad, bool clamp_on, const hot::Params& hp) {
    typedef hot::Lanes<T> L;
    if (grad == 0) {
        if (clamp_on) {
#pragma unroll
            for (int i = 0; i < L::NP; i++) L::set(v0, i, hot::fwd<ACT, true>(__fadd2_rn(L::get(v0, i), bv[i]), hp));
        } else {
#pragma unroll
            for (int i = 0; i < L::NP; i++) L::set(v0, i, hot::fwd<ACT, false>(__fadd2_rn(L::get(v0, i), bv[i]), hp));
        }
    } else {
        if (clamp_on) {
#pragma unroll
            for (int i = 0; i < L::NP; i++) L::set(v0, i, hot::bwd<ACT, true>(L::get(v0, i), L::get(yv, i), hp));
        } else {
#pragma unroll
            for (int i = 0; i < L::NP; i++) L::set(v0, i, hot::bwd<ACT, false>(L::get(v0, i), L::get(yv, i), hp));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Vector path for linear / lrelu (the only activations on the StyleGAN2 path).  BMODE: 0 = no bias, 1 = bias constant
// across a 16-byte vector (step_b % VEC == 0, e.g. NCHW), 2 = bias varies per element with step_b == 1 and
// size_b % VEC == 0 (channels-last tensors and [N, C] matrices).
// ---------------------------------------------------------------------------------------------------------------
template <class T, int ACT, int BMODE, bool USE_YREF>
__global__ void __launch_bounds__(256) bias_act_vec_kernel(BiasActParams p) {
    typedef typename Acc<T>::type S;
    constexpr int VEC = Vec16<T>::N;
    constexpr int UNROLL = 4;
    const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* yr = (const T*)p.yref; T* y = (T*)p.y;
    const long long nvec = p.size_x / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clampv = (S)p.clamp;
    const int grad = p.grad;
    const hot::Params hp = hot::make_params(p.alpha, p.gain, p.clamp);
    const bool clamp_on = p.clamp >= 0.f && (p.grad == 0 || USE_YREF);   // without a saved output the clamp cannot mask the gradient
    long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; iv < nvec; iv += stride * UNROLL) {
        Vec16<T> xv[UNROLL], yv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            long long j = iv + u * stride;
            if (j < nvec) {
                xv[u] = ld16_stream(x + j * VEC);
                if (USE_YREF) yv[u] = ld16_stream(yr + j * VEC);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            long long j = iv + u * stride;
            if (j < nvec) {
                long long e0 = j * VEC;
                float2 bf[hot::Lanes<T>::NP];
#pragma unroll
                for (int i = 0; i < hot::Lanes<T>::NP; i++) bf[i] = make_float2(0.f, 0.f);
                if (BMODE == 1 && grad == 0) {
                    const float bs = (float)to_acc<T>(__ldg(b + (e0 / p.step_b) % p.size_b));
#pragma unroll
                    for (int i = 0; i < hot::Lanes<T>::NP; i++) bf[i] = make_float2(bs, bs);
                }
                if (BMODE == 2 && grad == 0) {
                    const Vec16<T> bv = ld16(b + (e0 % p.size_b));
#pragma unroll
                    for (int i = 0; i < hot::Lanes<T>::NP; i++) bf[i] = hot::Lanes<T>::get(bv, i);
                }
                hot_vector<T, ACT>(xv[u], yv[u], bf, grad, clamp_on, hp);
                st16_stream(y + e0, xv[u]);
            }
        }
    }
    // ragged tail (size_x % VEC elements)
    if (blockIdx.x == 0) {
        long long i = nvec * VEC + threadIdx.x;
        if (i < p.size_x) {
            S v = to_acc<T>(x[i]);
            if (BMODE != 0 && grad == 0) v += to_acc<T>(b[(i / p.step_b) % p.size_b]);
            S yrv = USE_YREF ? to_acc<T>(yr[i]) : (S)0;
            y[i] = from_acc<T>(act_eval<ACT, S>(v, (S)0, yrv, (S)1, grad, alpha, gain, clampv));
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Bulk-staged vector path (stream_bulk.cuh): same arithmetic as bias_act_vec_kernel, data moved by the bulk copy engine.
// ---------------------------------------------------------------------------------------------------------------
template <class T, int ACT, int BMODE, bool USE_YREF>
__global__ void __launch_bounds__(streamk::NTHREADS) bias_act_bulk_kernel(BiasActParams p) {
    typedef typename Acc<T>::type S;
    constexpr int VEC = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* yr = (const T*)p.yref; T* y = (T*)p.y;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clampv = (S)p.clamp;
    const int grad = p.grad;
    const hot::Params hp = hot::make_params(p.alpha, p.gain, p.clamp);
    const bool clamp_on = p.clamp >= 0.f && (p.grad == 0 || USE_YREF);   // without a saved output the clamp cannot mask the gradient
    typedef hot::Lanes<T> L;
    // channels-last bias: a thread's vectors always cover the same channels when size_b divides 256 * VEC -> fetch once
    const bool fixed_b = BMODE == 2 && grad == 0 && (256 * VEC) % p.size_b == 0;
    float2 bfix[L::NP];
#pragma unroll
    for (int i = 0; i < L::NP; i++) bfix[i] = make_float2(0.f, 0.f);
    if (fixed_b) {
        const Vec16<T> bv = ld16(b + ((int)threadIdx.x * VEC) % p.size_b);
#pragma unroll
        for (int i = 0; i < L::NP; i++) bfix[i] = L::get(bv, i);
    }
    auto body = [&](long long e0, Vec16<T>& v0, const Vec16<T>& v1, const Vec16<T>&, bool live, uint32_t, int) {
        if (!live) return;
        if (BMODE == 0 || grad != 0 || fixed_b) {
            hot_vector<T, ACT>(v0, v1, bfix, grad, clamp_on, hp);
        } else {
            float2 bf[L::NP];
            if (BMODE == 1) {
                const float bs = (float)to_acc<T>(__ldg(b + (e0 / p.step_b) % p.size_b));
#pragma unroll
                for (int i = 0; i < L::NP; i++) bf[i] = make_float2(bs, bs);
            } else {
                const Vec16<T> bv = ld16(b + (e0 % p.size_b));
#pragma unroll
                for (int i = 0; i < L::NP; i++) bf[i] = L::get(bv, i);
            }
            hot_vector<T, ACT>(v0, v1, bf, grad, clamp_on, hp);
        }
    };
    if (USE_YREF) streamk::run<T, 2, 6, 8192>(x, yr, (const T*)nullptr, y, p.size_x, smem_raw, body);
    else streamk::run<T, 1, 6, 16384>(x, (const T*)nullptr, (const T*)nullptr, y, p.size_x, smem_raw, body);
    // ragged tail (size_x % VEC elements)
    if (blockIdx.x == 0) {
        long long i = (p.size_x / VEC) * VEC + threadIdx.x;
        if (i < p.size_x) {
            S v = to_acc<T>(x[i]);
            if (BMODE != 0 && grad == 0) v += to_acc<T>(b[(i / p.step_b) % p.size_b]);
            S yrv = USE_YREF ? to_acc<T>(yr[i]) : (S)0;
            y[i] = from_acc<T>(act_eval<ACT, S>(v, (S)0, yrv, (S)1, grad, alpha, gain, clampv));
        }
    }
}

template <class K>
int bulk_smem_attr(K kernel, int bytes, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        gt_set_error("%s: cannot reserve %d bytes of shared memory: %s", name, bytes, cudaGetErrorString(e));
        return GT_ERR_CUDA;
    }
    return GT_OK;
}

template <class T, int ACT, int BMODE, bool USE_YREF>
int launch_bulk_t(const BiasActParams& p, cudaStream_t st) {
    constexpr int SMEM = USE_YREF ? streamk::Smem<2, 6, 8192>::TOTAL : streamk::Smem<1, 6, 16384>::TOTAL;
    static bool configured = false;
    if (!configured) {
        int rc = bulk_smem_attr(bias_act_bulk_kernel<T, ACT, BMODE, USE_YREF>, SMEM, "gt_bias_act(bulk)");
        if (rc != GT_OK) return rc;
        configured = true;
    }
    const int grid = streamk::grid_for(p.size_x, (int)sizeof(T), USE_YREF ? 8192 : 16384, 2);
    bias_act_bulk_kernel<T, ACT, BMODE, USE_YREF><<<grid, streamk::NTHREADS, SMEM, st>>>(p);
    GT_CUDA_LAUNCH_CHECK("gt_bias_act(bulk)");
    return GT_OK;
}

template <class T, int ACT>
int launch_scalar(const BiasActParams& p, cudaStream_t st) {
    long long blocks = (p.size_x + 255) / 256;
    long long cap = (long long)gt_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    bias_act_scalar_kernel<T, ACT><<<(unsigned)blocks, 256, 0, st>>>(p);
    GT_CUDA_LAUNCH_CHECK("gt_bias_act(scalar)");
    return GT_OK;
}

template <class T>
int dispatch_scalar(const BiasActParams& p, int act, cudaStream_t st) {
    switch (act) {
        case A_LINEAR: return launch_scalar<T, A_LINEAR>(p, st);
        case A_RELU: return launch_scalar<T, A_RELU>(p, st);
        case A_LRELU: return launch_scalar<T, A_LRELU>(p, st);
        case A_TANH: return launch_scalar<T, A_TANH>(p, st);
        case A_SIGMOID: return launch_scalar<T, A_SIGMOID>(p, st);
        case A_ELU: return launch_scalar<T, A_ELU>(p, st);
        case A_SELU: return launch_scalar<T, A_SELU>(p, st);
        case A_SOFTPLUS: return launch_scalar<T, A_SOFTPLUS>(p, st);
        case A_SWISH: return launch_scalar<T, A_SWISH>(p, st);
    }
    gt_set_error("gt_bias_act: unknown activation id %d", act);
    return GT_ERR_ARG;
}

template <class T, int ACT, int BMODE>
int launch_vec(const BiasActParams& p, cudaStream_t st) {
    constexpr int VEC = Vec16<T>::N;
    long long nvec = p.size_x / VEC;
    if (gt_stream_variant() == 0 && p.size_x * (long long)sizeof(T) >= (1ll << 22)) {   // >= 4 MB: bulk-staged streaming
        if (p.yref) return launch_bulk_t<T, ACT, BMODE, true>(p, st);
        return launch_bulk_t<T, ACT, BMODE, false>(p, st);
    }
    // 4 vectors per thread per trip; whole waves of 8 resident 256-thread CTAs per SM.
    long long blocks = (nvec + 256 * 4 - 1) / (256 * 4);
    long long wave = (long long)gt_num_sms() * 8;
    if (blocks > wave) blocks = wave;
    if (blocks < 1) blocks = 1;
    if (p.yref)
        bias_act_vec_kernel<T, ACT, BMODE, true><<<(unsigned)blocks, 256, 0, st>>>(p);
    else
        bias_act_vec_kernel<T, ACT, BMODE, false><<<(unsigned)blocks, 256, 0, st>>>(p);
    GT_CUDA_LAUNCH_CHECK("gt_bias_act(vec)");
    return GT_OK;
}

inline bool aligned16(const void* q) { return q == nullptr || (((uintptr_t)q) & 15) == 0; }

template <class T>
int dispatch(const BiasActParams& p, int act, cudaStream_t st) {
    constexpr int VEC = Vec16<T>::N;
    bool hot = (act == A_LINEAR || act == A_LRELU) && p.grad <= 1 && p.xref == nullptr && p.dy == nullptr;
    bool al = aligned16(p.x) && aligned16(p.y) && aligned16(p.yref) && aligned16(p.b);
    if (hot && al && p.size_x >= VEC) {
        int bmode = -1;
        if (p.b == nullptr || p.grad != 0) bmode = 0;   // bias only enters the forward pass for linear / lrelu
        else if (p.step_b % VEC == 0) bmode = 1;
        else if (p.step_b == 1 && p.size_b % VEC == 0) bmode = 2;
        if (bmode >= 0) {
            if (act == A_LINEAR) {
                if (bmode == 0) return launch_vec<T, A_LINEAR, 0>(p, st);
                if (bmode == 1) return launch_vec<T, A_LINEAR, 1>(p, st);
                return launch_vec<T, A_LINEAR, 2>(p, st);
            } else {
                if (bmode == 0) return launch_vec<T, A_LRELU, 0>(p, st);
                if (bmode == 1) return launch_vec<T, A_LRELU, 1>(p, st);
                return launch_vec<T, A_LRELU, 2>(p, st);
            }
        }
    }
    return dispatch_scalar<T>(p, act, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Fused backward for linear / lrelu: dx = grad-1 pass of dy, plus deterministic per-channel partial sums of dx.
// Layout: x viewed as [outer, size_b, inner] (NCHW: outer=N, inner=H*W; channels-last / [N,C]: inner=1).
// Each CTA owns one (channel, slice) pair and writes one partial: db_partial[slice * size_b + channel]; a second
// tiny kernel adds the partials in a fixed order, so the result does not depend on scheduling (replicas must stay
// bit-identical: S3/torch_utils/misc.py:180-191).
// ---------------------------------------------------------------------------------------------------------------
template <class T, int ACT>
__global__ void __launch_bounds__(256) bias_act_bwd_plane_kernel(const T* __restrict__ dy, const T* __restrict__ yref, T* __restrict__ dx,
                                                                 float* __restrict__ partial, int C, long long inner, int outer,
                                                                 int slices, float alpha, float gain, float clampv) {
    // grid.x = C * slices ; each CTA walks n in [slice, outer) step slices over its channel plane.
    constexpr int VEC = Vec16<T>::N;
    const int c = blockIdx.x % C;
    const int slice = blockIdx.x / C;
    const hot::Params hp = hot::make_params(alpha, gain, clampv);
    float acc = 0.f;
    const bool vec_ok = (inner % VEC == 0);
    for (int n = slice; n < outer; n += slices) {
        const long long base = ((long long)n * C + c) * inner;
        if (vec_ok) {
            for (long long i = (long long)threadIdx.x * VEC; i < inner; i += 256 * VEC) {
                Vec16<T> g = ld16_stream(dy + base + i), yv, o;
                if (yref) yv = ld16_stream(yref + base + i);
                if (!yref) yv = g;
                o = g;
                hot_vector<T, ACT>(o, yv, nullptr, 1, yref != nullptr && clampv >= 0.f, hp);
#pragma unroll
                for (int k = 0; k < hot::Lanes<T>::NP; k++) {
                    const float2 st = hot::Lanes<T>::stored(o, k);   // sum what is stored, like dx.sum() on the stored tensor
                    acc += st.x;
                    acc += st.y;
                }
                st16_stream(dx + base + i, o);
            }
        } else {
            for (long long i = threadIdx.x; i < inner; i += 256) {
                float r = act_eval<ACT, float>((float)to_acc<T>(dy[base + i]), 0.f, yref ? (float)to_acc<T>(yref[base + i]) : 0.f, 1.f, 1, alpha, gain, clampv);
                T o = from_acc<T>(r);
                dx[base + i] = o;
                acc += (float)to_acc<T>(o);
            }
        }
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) s += red[w];
        partial[(long long)slice * C + c] = s;
    }
}

// channels-last / matrix layout: element index = row * C + c (C % VEC == 0).  Each CTA owns a band of rows.  A CTA is
// split into row groups of `lanes` threads; thread (rg, cv) walks rows band + bands*(rg + rgroups*k) for channel
// vector cv, so a warp always reads consecutive 16-byte vectors.  Row groups are combined through shared memory in
// a fixed order.  Partials: partial[band * C + c].
template <class T, int ACT>
__global__ void __launch_bounds__(256) bias_act_bwd_cl_kernel(const T* __restrict__ dy, const T* __restrict__ yref, T* __restrict__ dx,
                                                              float* __restrict__ partial, int C, long long rows, int bands,
                                                              float alpha, float gain, float clampv) {
    constexpr int VEC = Vec16<T>::N;
    __shared__ float red[256 * VEC];
    const hot::Params hp = hot::make_params(alpha, gain, clampv);
    const int band = blockIdx.x;
    const int cvecs = C / VEC;
    const int lanes = cvecs < 256 ? cvecs : 256;
    const int rgroups = 256 / lanes;
    const int rg = threadIdx.x / lanes;
    const int lane = threadIdx.x - rg * lanes;
    for (int cv0 = 0; cv0 < cvecs; cv0 += lanes) {
        const int cv = cv0 + lane;
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[k] = 0.f;
        if (rg < rgroups && cv < cvecs) {
            for (long long r = band + (long long)bands * rg; r < rows; r += (long long)bands * rgroups) {
                long long e0 = r * C + (long long)cv * VEC;
                Vec16<T> g = ld16_stream(dy + e0), yv, o;
                if (yref) yv = ld16_stream(yref + e0);
                if (!yref) yv = g;
                o = g;
                hot_vector<T, ACT>(o, yv, nullptr, 1, yref != nullptr && clampv >= 0.f, hp);
#pragma unroll
                for (int k = 0; k < hot::Lanes<T>::NP; k++) {
                    const float2 st = hot::Lanes<T>::stored(o, k);
                    acc[2 * k] += st.x;
                    acc[2 * k + 1] += st.y;
                }
                st16_stream(dx + e0, o);
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) red[threadIdx.x * VEC + k] = acc[k];
        __syncthreads();
        if (rg == 0 && cv < cvecs) {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                float s = 0.f;
                for (int g2 = 0; g2 < rgroups; g2++) s += red[(g2 * lanes + lane) * VEC + k];
                partial[(long long)band * C + cv * VEC + k] = s;
            }
        }
        __syncthreads();
    }
}


// Bulk-staged variant of bias_act_bwd_cl_kernel.  A thread's vectors always cover the same VEC channels (see
// stream_bulk.cuh), so the bias-gradient partial sums live in registers; they are combined across the CTA through shared
// memory in a fixed order.  Partials: partial[blockIdx.x * C + c].
template <class T, int ACT, bool USE_YREF>
__global__ void __launch_bounds__(streamk::NTHREADS) bias_act_bwd_bulk_kernel(const T* __restrict__ dy, const T* __restrict__ yref, T* __restrict__ dx,
                                                                float* __restrict__ partial, int C, long long nelem, float alpha, float gain,
                                                                float clampv) {
    constexpr int VEC = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    __shared__ float red[256 * VEC];
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) acc[k] = 0.f;
    const hot::Params hp = hot::make_params(alpha, gain, clampv);
    const bool clamp_on = USE_YREF && clampv >= 0.f;
    auto body = [&](long long, Vec16<T>& v0, const Vec16<T>& v1, const Vec16<T>&, bool live, uint32_t, int) {
        if (!live) return;
        hot_vector<T, ACT>(v0, USE_YREF ? v1 : v0, nullptr, 1, clamp_on, hp);
#pragma unroll
        for (int k = 0; k < hot::Lanes<T>::NP; k++) {
            const float2 st = hot::Lanes<T>::stored(v0, k);   // sum what is stored, like dx.sum() on the stored tensor
            acc[2 * k] += st.x;
            acc[2 * k + 1] += st.y;
        }
    };
    if (USE_YREF) streamk::run<T, 2, 6, 8192>(dy, yref, (const T*)nullptr, dx, nelem, smem_raw, body);
    else streamk::run<T, 1, 6, 16384>(dy, (const T*)nullptr, (const T*)nullptr, dx, nelem, smem_raw, body);
    if (threadIdx.x < 256) {              // (the copy-engine driver warp holds no sums)
#pragma unroll
        for (int k = 0; k < VEC; k++) red[threadIdx.x * VEC + k] = acc[k];
    }
    __syncthreads();
    const int cvecs = C / VEC;
    if ((int)threadIdx.x < cvecs) {
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            float s = 0.f;
            for (int g2 = threadIdx.x; g2 < 256; g2 += cvecs) s += red[g2 * VEC + k];
            partial[(long long)blockIdx.x * C + threadIdx.x * VEC + k] = s;
        }
    }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ db, int C, int parts) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int j = 0; j < parts; j++) s += partial[(long long)j * C + c];
    db[c] = s;
}

}  // namespace

extern "C" int gt_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y, int dtype,
                           int grad, int act, float alpha, float gain, float clamp, long long size_x, int size_b,
                           long long step_b, void* stream) {
    GT_REQUIRE(x && y, "gt_bias_act: x and y must be non-null");
    GT_REQUIRE(size_x >= 0, "gt_bias_act: negative size");
    GT_REQUIRE(grad >= 0 && grad <= 2, "gt_bias_act: grad must be 0, 1 or 2 (got %d)", grad);
    GT_REQUIRE(b == nullptr || (size_b > 0 && step_b > 0), "gt_bias_act: bias given with size_b=%d step_b=%lld", size_b, step_b);
    if (size_x == 0) return GT_OK;
    BiasActParams p;
    p.x = x; p.b = b; p.xref = xref; p.yref = yref; p.dy = dy; p.y = y;
    p.grad = grad; p.alpha = alpha; p.gain = gain; p.clamp = clamp;
    p.size_x = size_x; p.size_b = b ? size_b : 1; p.step_b = b ? step_b : 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case GT_F32: return dispatch<float>(p, act, st);
        case GT_F16: return dispatch<__half>(p, act, st);
        case GT_F64: return dispatch_scalar<double>(p, act, st);
    }
    gt_set_error("gt_bias_act: unsupported dtype code %d", dtype);
    return GT_ERR_ARG;
}

template <class T>
static int bias_act_bwd_t(const T* dy, const T* yref, T* dx, float* db, float* ws, long long ws_floats, int act, float alpha, float gain,
                          float clamp, int outer, int C, long long inner, cudaStream_t st) {
    constexpr int VEC = Vec16<T>::N;
    const int sms = gt_num_sms();
    int parts;
    if (inner > 1 || C % VEC != 0) {
        // plane layout (also the generic fallback): pick slices so that C*slices ~ 4 waves of CTAs
        int slices = (sms * 8 + C - 1) / C;
        if (slices > outer) slices = outer;
        if (slices < 1) slices = 1;
        parts = slices;
        GT_REQUIRE((long long)parts * C <= ws_floats, "gt_bias_act_bwd: workspace too small (%lld floats, need %lld)", ws_floats, (long long)parts * C);
        if (act == A_LINEAR)
            bias_act_bwd_plane_kernel<T, A_LINEAR><<<C * slices, 256, 0, st>>>(dy, yref, dx, ws, C, inner, outer, slices, alpha, gain, clamp);
        else
            bias_act_bwd_plane_kernel<T, A_LRELU><<<C * slices, 256, 0, st>>>(dy, yref, dx, ws, C, inner, outer, slices, alpha, gain, clamp);
    } else {
        long long rows = outer;
        const long long nelem = rows * C;
        const bool al = ((((uintptr_t)dy) | ((uintptr_t)dx) | ((uintptr_t)yref)) & 15) == 0;
        if (gt_stream_variant() == 0 && al && (256 * VEC) % C == 0 && nelem * (long long)sizeof(T) >= (1ll << 22)) {
            const int ch = yref ? 8192 : 16384;
            const int grid = streamk::grid_for(nelem, (int)sizeof(T), ch, 2);
            parts = grid;
            GT_REQUIRE((long long)parts * C <= ws_floats, "gt_bias_act_bwd: workspace too small (%lld floats, need %lld)", ws_floats, (long long)parts * C);
#define GT_BWD_BULK(ACT_, YR_)                                                                                                         \
    {                                                                                                                                   \
        constexpr int SMEM = YR_ ? streamk::Smem<2, 6, 8192>::TOTAL : streamk::Smem<1, 6, 16384>::TOTAL;                                  \
        static bool configured = false;                                                                                                 \
        if (!configured) {                                                                                                              \
            int rc = bulk_smem_attr(bias_act_bwd_bulk_kernel<T, ACT_, YR_>, SMEM, "gt_bias_act_bwd(bulk)");                              \
            if (rc != GT_OK) return rc;                                                                                                 \
            configured = true;                                                                                                          \
        }                                                                                                                               \
        bias_act_bwd_bulk_kernel<T, ACT_, YR_><<<grid, streamk::NTHREADS, SMEM, st>>>(dy, yref, dx, ws, C, nelem, alpha, gain, clamp);                 \
    }
            if (act == A_LINEAR) {
                if (yref) GT_BWD_BULK(A_LINEAR, true) else GT_BWD_BULK(A_LINEAR, false)
            } else {
                if (yref) GT_BWD_BULK(A_LRELU, true) else GT_BWD_BULK(A_LRELU, false)
            }
#undef GT_BWD_BULK
            GT_CUDA_LAUNCH_CHECK("gt_bias_act_bwd(bulk)");
            reduce_partials_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws, db, C, parts);
            GT_CUDA_LAUNCH_CHECK("gt_bias_act_bwd(reduce)");
            return GT_OK;
        }
        int bands = sms * 4;
        if (bands > rows) bands = (int)rows;
        if (bands < 1) bands = 1;
        parts = bands;
        GT_REQUIRE((long long)parts * C <= ws_floats, "gt_bias_act_bwd: workspace too small (%lld floats, need %lld)", ws_floats, (long long)parts * C);
        if (act == A_LINEAR)
            bias_act_bwd_cl_kernel<T, A_LINEAR><<<bands, 256, 0, st>>>(dy, yref, dx, ws, C, rows, bands, alpha, gain, clamp);
        else
            bias_act_bwd_cl_kernel<T, A_LRELU><<<bands, 256, 0, st>>>(dy, yref, dx, ws, C, rows, bands, alpha, gain, clamp);
    }
    GT_CUDA_LAUNCH_CHECK("gt_bias_act_bwd");
    reduce_partials_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws, db, C, parts);
    GT_CUDA_LAUNCH_CHECK("gt_bias_act_bwd(reduce)");
    return GT_OK;
}

extern "C" long long gt_bias_act_bwd_workspace(int outer, int C, long long inner) {
    int sms = gt_num_sms();
    long long a = (long long)((sms * 8 + C - 1) / C + 1) * C;
    long long b = (long long)sms * 4 * C;
    (void)outer; (void)inner;
    return a > b ? a : b;
}

extern "C" int gt_bias_act_bwd(const void* dy, const void* yref, void* dx, float* db, float* workspace, long long workspace_floats,
                               int dtype, int act, float alpha, float gain, float clamp, int outer, int C, long long inner, void* stream) {
    GT_REQUIRE(dy && dx && db && workspace, "gt_bias_act_bwd: null pointer");
    GT_REQUIRE(act == A_LINEAR || act == A_LRELU, "gt_bias_act_bwd: only linear (1) and lrelu (3) are fused; got %d", act);
    GT_REQUIRE(outer > 0 && C > 0 && inner > 0, "gt_bias_act_bwd: empty shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GT_F32) return bias_act_bwd_t<float>((const float*)dy, (const float*)yref, (float*)dx, db, workspace, workspace_floats, act, alpha, gain, clamp, outer, C, inner, st);
    if (dtype == GT_F16) return bias_act_bwd_t<__half>((const __half*)dy, (const __half*)yref, (__half*)dx, db, workspace, workspace_floats, act, alpha, gain, clamp, outer, C, inner, st);
    gt_set_error("gt_bias_act_bwd: unsupported dtype code %d", dtype);
    return GT_ERR_ARG;
}
