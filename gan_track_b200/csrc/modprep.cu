// modprep.cu -- the parameter / style side of the training-mode modulated convolution in a handful of launches.
//
// Reference: modulated_conv2d (S3/training/networks_stylegan2.py:52-63; S3 = /root/reference/src/models/stylegan3)
//     [fp16]  weight = weight * (1 / sqrt(I kh kw) / weight.norm(inf, dim=[1,2,3]))      styles = styles / styles.norm(inf, dim=1)
//             dcoefs = rsqrt(sum_{i,k} (weight[o,i,k] styles[n,i])^2 + 1e-8)  =  rsqrt(styles^2 @ wsq^T + 1e-8),  wsq[o,i] = sum_k weight[o,i,k]^2
// which autograd runs as ~13 tiny kernels forward and ~20 backward per layer (max-reduce, reciprocal, broadcast multiply,
// square, sum, cast, ...): ~2.5 ms of a 53 ms training iteration.  Here:
//     modprep_weight_fwd   one CTA per output channel: max |w|, scale, cast, wsq                (w side)
//     modprep_style_fwd    one CTA per sample: max |s|, sn = s / max, sn^2                      (style side)
//     [gt_fc_fwd]          q = sn^2 @ wsq^T                                                      (csrc/fc.cu)
//     modprep_rsqrt        d = rsqrt(q + eps)
// and the vector-Jacobian products
//     modprep_gq           gq = -1/2 d^3 gd
//     [gt_fc_wgrad]        g_wsq = gq^T @ sn^2          [gt_fc_dgrad]  t = gq @ wsq
//     modprep_style_bwd    g_sn += 2 sn t;  gs = g_sn / m - [i = argmax] sign(s_i) (g_sn . sn) / m
//     modprep_weight_bwd   g_wn = g_w + 2 wn g_wsq;  gW = a g_wn - [j = argmax] sign(W_j) (g_wn . W) a / m
// (a = c / m the per-channel scale; without pre-normalisation a = 1 and the max terms vanish).  The max is taken at its
// first occurrence; ties have measure zero for float weights / styles.
//
// Second order (the path-length regulariser differentiates the generator's backward, S3/training/loss.py:85-100).  The weight side is
// never differentiated twice there (it is not on a path to the latents).  The style-side vector-Jacobian product
//     (g_sn, g_d, s, wsq) -> gs          [with gq = -1/2 d^3 g_d,  gp = gq @ wsq,  tt = g_sn + 2 sn gp]
// has, for a cotangent u of gs, the closed-form derivative (r = sign(s_k) u_k at the arg-max k, 0 without pre-normalisation; m = max|s|)
//     v = u - r sn                      z = (v sn) @ wsq^T
//     d/d g_sn = v / m                  d/d g_d = -d^3 z / m
//     hq = 3/2 d^5 g_d z / m            hp = hq @ wsq
//     D  = (2 gp v - r tt) / m + 2 sn hp
//     d/d s = D / m + [i = k] sign(s_k) (-(tt . v) / m^2 - (D . sn) / m)
//     d/d wsq = gq^T @ (2 v sn / m) + hq^T @ sn^2
// evaluated by three small kernels (style_bwd2_a / _b / _c) around three calls of the fully-connected kernels (csrc/fc.cu) --
// ~10 launches per layer instead of the ~70 autograd needs for the op-by-op chain.  Checked against autograd's double backward of the
// tensor-op form in tests/test_gpu_modulated.py (and in float64 on the CPU: tests/test_modprep_closed_form.py).
#include "gt_common.cuh"

#include <cuda_fp16.h>

namespace {

struct MaxIdx {
    float v;
    int i;
};

__device__ __forceinline__ MaxIdx max_first(MaxIdx a, MaxIdx b) {
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

__device__ __forceinline__ MaxIdx block_max_first(MaxIdx m, MaxIdx* sh) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        MaxIdx o;
        o.v = __shfl_xor_sync(0xffffffffu, m.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, m.i, off);
        m = max_first(m, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = m;
    __syncthreads();
    MaxIdx r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = max_first(r, sh[w]);
    return r;
}

__device__ __forceinline__ float block_sum_fixed(float v, float* sh) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += sh[w];
    return r;
}

// grid = O.  W [O, I, KK] fp32 -> w16 [O, I, KK] fp16 (prenorm only), wsq [O, I], scale [O] (a_o), amax [O] (argmax within the row)
__global__ void __launch_bounds__(256) modprep_weight_fwd_kernel(const float* __restrict__ W, __half* __restrict__ w16, float* __restrict__ wsq,
                                                                  float* __restrict__ scale, int* __restrict__ amax, int I, int KK, int prenorm, float c) {
    __shared__ MaxIdx shm[8];
    const int o = blockIdx.x, L = I * KK;
    const float* row = W + (long long)o * L;
    float a = 1.f;
    if (prenorm) {
        MaxIdx m;
        m.v = -1.f;
        m.i = 0x7fffffff;
        for (int j = threadIdx.x; j < L; j += 256) {
            MaxIdx e;
            e.v = fabsf(row[j]);
            e.i = j;
            m = max_first(m, e);
        }
        m = block_max_first(m, shm);
        a = __fmul_rn(__fdiv_rn(1.f, m.v), c);      // the reference's scalar / tensor is reciprocal(tensor) * scalar (Tensor.__rtruediv__)
        if (threadIdx.x == 0) {
            scale[o] = a;
            amax[o] = m.i;
        }
    }
    for (int i = threadIdx.x; i < I; i += 256) {
        float s = 0.f;
        for (int k = 0; k < KK; k++) {
            const float wn = row[i * KK + k] * a;
            s += wn * wn;
            if (prenorm) w16[(long long)o * L + i * KK + k] = __float2half_rn(wn);
        }
        wsq[(long long)o * I + i] = s;
    }
}

// grid = N.  s [N, I] -> sn [N, I], sn2 [N, I], smax [N], sarg [N]
__global__ void __launch_bounds__(256) modprep_style_fwd_kernel(const float* __restrict__ s, float* __restrict__ sn, float* __restrict__ sn2,
                                                                 float* __restrict__ smax, int* __restrict__ sarg, int I, int prenorm) {
    __shared__ MaxIdx shm[8];
    const int n = blockIdx.x;
    const float* row = s + (long long)n * I;
    float m = 1.f;
    if (prenorm) {
        MaxIdx mi;
        mi.v = -1.f;
        mi.i = 0x7fffffff;
        for (int j = threadIdx.x; j < I; j += 256) {
            MaxIdx e;
            e.v = fabsf(row[j]);
            e.i = j;
            mi = max_first(mi, e);
        }
        mi = block_max_first(mi, shm);
        m = mi.v;
        if (threadIdx.x == 0) {
            smax[n] = m;
            sarg[n] = mi.i;
        }
    }
    for (int i = threadIdx.x; i < I; i += 256) {
        const float v = prenorm ? row[i] / m : row[i];
        if (prenorm) sn[(long long)n * I + i] = v;
        sn2[(long long)n * I + i] = v * v;
    }
}

__global__ void modprep_rsqrt_kernel(const float* __restrict__ q, float* __restrict__ d, int n, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = rsqrtf(q[i] + eps);
}

__global__ void modprep_gq_kernel(const float* __restrict__ d, const float* __restrict__ gd, float* __restrict__ gq, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float dd = d[i];
        gq[i] = -0.5f * dd * dd * dd * gd[i];
    }
}

// grid = N.  gs = d(loss)/d(styles) from g_sn (gradient w.r.t. the normalised styles, may be null) and t = gq @ wsq
__global__ void __launch_bounds__(256) modprep_style_bwd_kernel(const float* __restrict__ g_sn, const float* __restrict__ sn, const float* __restrict__ t,
                                                                 const float* __restrict__ smax, const int* __restrict__ sarg, float* __restrict__ gs, int I,
                                                                 int prenorm) {
    __shared__ float shs[8];
    const int n = blockIdx.x;
    const long long base = (long long)n * I;
    float dot = 0.f;
    for (int i = threadIdx.x; i < I; i += 256) {
        const float g = (g_sn ? g_sn[base + i] : 0.f) + 2.f * sn[base + i] * t[base + i];
        dot += g * sn[base + i];
    }
    if (!prenorm) {
        for (int i = threadIdx.x; i < I; i += 256) gs[base + i] = (g_sn ? g_sn[base + i] : 0.f) + 2.f * sn[base + i] * t[base + i];
        return;
    }
    dot = block_sum_fixed(dot, shs);
    const float m = smax[n];
    const int j = sarg[n];
    for (int i = threadIdx.x; i < I; i += 256) {
        const float g = (g_sn ? g_sn[base + i] : 0.f) + 2.f * sn[base + i] * t[base + i];
        float r = g / m;
        if (i == j) r -= (sn[base + i] >= 0.f ? 1.f : -1.f) * dot / m;
        gs[base + i] = r;
    }
}

// grid = O.  gW from g_w (gradient w.r.t. the scaled [and cast] weight, may be null) and g_wsq
template <class TG>
__global__ void __launch_bounds__(256) modprep_weight_bwd_kernel(const float* __restrict__ W, const TG* __restrict__ g_w, const float* __restrict__ g_wsq,
                                                                  const float* __restrict__ scale, const int* __restrict__ amax, float* __restrict__ gW, int I,
                                                                  int KK, int prenorm) {
    __shared__ float shs[8];
    const int o = blockIdx.x, L = I * KK;
    const long long base = (long long)o * L;
    const float a = prenorm ? scale[o] : 1.f;
    float dot = 0.f;
    if (prenorm) {
        for (int j = threadIdx.x; j < L; j += 256) {
            const float w = W[base + j];
            const float g = (g_w ? (float)g_w[base + j] : 0.f) + 2.f * a * w * g_wsq[(long long)o * I + j / KK];
            dot += g * w;
        }
        dot = block_sum_fixed(dot, shs);
    }
    const int jm = prenorm ? amax[o] : -1;
    const float wm = prenorm ? W[base + jm] : 1.f;
    for (int j = threadIdx.x; j < L; j += 256) {
        const float w = W[base + j];
        const float g = (g_w ? (float)g_w[base + j] : 0.f) + 2.f * a * w * g_wsq[(long long)o * I + j / KK];
        float r = a * g;
        if (j == jm) r -= (wm >= 0.f ? 1.f : -1.f) * dot * a / fabsf(wm);
        gW[base + j] = r;
    }
}


// ---- second order of the style side (see the header) ------------------------------------------------------------------------------

// grid = N.  u [N, I] -> v = u - r sn, vs = v sn, r [N]
__global__ void __launch_bounds__(256) modprep_style_bwd2_a_kernel(const float* __restrict__ u, const float* __restrict__ sn, const int* __restrict__ sarg,
                                                                    float* __restrict__ v, float* __restrict__ vs, float* __restrict__ r_out, int I, int prenorm) {
    const int n = blockIdx.x;
    const long long base = (long long)n * I;
    float r = 0.f;
    if (prenorm) {
        const int k = sarg[n];
        r = (sn[base + k] >= 0.f ? 1.f : -1.f) * u[base + k];
    }
    if (threadIdx.x == 0) r_out[n] = r;
    for (int i = threadIdx.x; i < I; i += 256) {
        const float x = sn[base + i];
        const float vv = u[base + i] - r * x;
        v[base + i] = vv;
        vs[base + i] = vv * x;
    }
}

// element-wise over [N, O]: ggd = -d^3 z / m, gq = -1/2 d^3 gd, hq = 3/2 d^5 gd z / m
__global__ void modprep_style_bwd2_b_kernel(const float* __restrict__ d, const float* __restrict__ gd, const float* __restrict__ z, const float* __restrict__ smax,
                                            float* __restrict__ ggd, float* __restrict__ gq, float* __restrict__ hq, int total, int O, int prenorm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float im = prenorm ? 1.f / smax[i / O] : 1.f;
    const float dd = d[i], d3 = dd * dd * dd, g = gd[i], zz = z[i] * im;
    ggd[i] = -d3 * zz;
    gq[i] = -0.5f * d3 * g;
    hq[i] = 1.5f * d3 * dd * dd * g * zz;
}

// grid = N.  -> gga = v / m (optional), g2s, x1 = 2 v sn / m, p = sn^2
__global__ void __launch_bounds__(256) modprep_style_bwd2_c_kernel(const float* __restrict__ a, const float* __restrict__ sn, const float* __restrict__ gp,
                                                                    const float* __restrict__ hp, const float* __restrict__ v, const float* __restrict__ r_in,
                                                                    const float* __restrict__ smax, const int* __restrict__ sarg, float* __restrict__ gga,
                                                                    float* __restrict__ g2s, float* __restrict__ x1, float* __restrict__ p_out, int I,
                                                                    int prenorm) {
    __shared__ float shs[8];
    const int n = blockIdx.x;
    const long long base = (long long)n * I;
    const float m = prenorm ? smax[n] : 1.f, im = 1.f / m, r = r_in[n];
    float L = 0.f, S2 = 0.f;
    for (int i = threadIdx.x; i < I; i += 256) {
        const float x = sn[base + i], g = gp[base + i], vv = v[base + i];
        const float tt = (a ? a[base + i] : 0.f) + 2.f * x * g;
        const float D = (2.f * g * vv - r * tt) * im + 2.f * x * hp[base + i];
        L += tt * vv;
        S2 += D * x;
    }
    if (prenorm) {
        L = block_sum_fixed(L, shs);
        S2 = block_sum_fixed(S2, shs);
    }
    const int k = prenorm ? sarg[n] : -1;
    const float dm = -(L * im) * im - S2 * im;
    for (int i = threadIdx.x; i < I; i += 256) {
        const float x = sn[base + i], g = gp[base + i], vv = v[base + i];
        const float tt = (a ? a[base + i] : 0.f) + 2.f * x * g;
        const float D = (2.f * g * vv - r * tt) * im + 2.f * x * hp[base + i];
        float out = D * im;
        if (i == k) out += (x >= 0.f ? 1.f : -1.f) * dm;
        g2s[base + i] = out;
        if (gga) gga[base + i] = vv * im;
        x1[base + i] = 2.f * vv * x * im;
        p_out[base + i] = x * x;
    }
}

}  // namespace

extern "C" int gt_modprep_weight_fwd(const float* W, void* w16, float* wsq, float* scale, int* amax, int O, int I, int KK, int prenorm, void* stream) {
    GT_REQUIRE(W && wsq, "gt_modprep_weight_fwd: null pointer");
    GT_REQUIRE(O > 0 && I > 0 && KK > 0, "gt_modprep_weight_fwd: empty weight");
    GT_REQUIRE(!prenorm || (w16 && scale && amax), "gt_modprep_weight_fwd: pre-normalisation needs w16, scale and amax outputs");
    const float c = (float)(1.0 / sqrt((double)I * KK));
    modprep_weight_fwd_kernel<<<O, 256, 0, (cudaStream_t)stream>>>(W, (__half*)w16, wsq, scale, amax, I, KK, prenorm ? 1 : 0, c);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_weight_fwd");
    return GT_OK;
}

extern "C" int gt_modprep_style_fwd(const float* s, float* sn, float* sn2, float* smax, int* sarg, int N, int I, int prenorm, void* stream) {
    GT_REQUIRE(s && sn2, "gt_modprep_style_fwd: null pointer");
    GT_REQUIRE(N > 0 && I > 0, "gt_modprep_style_fwd: empty styles");
    GT_REQUIRE(!prenorm || (sn && smax && sarg), "gt_modprep_style_fwd: pre-normalisation needs sn, smax and sarg outputs");
    modprep_style_fwd_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(s, sn, sn2, smax, sarg, I, prenorm ? 1 : 0);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_style_fwd");
    return GT_OK;
}

extern "C" int gt_modprep_rsqrt(const float* q, float* d, int n, float eps, void* stream) {
    GT_REQUIRE(q && d && n > 0, "gt_modprep_rsqrt: null pointer or empty");
    modprep_rsqrt_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(q, d, n, eps);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_rsqrt");
    return GT_OK;
}

extern "C" int gt_modprep_gq(const float* d, const float* gd, float* gq, int n, void* stream) {
    GT_REQUIRE(d && gd && gq && n > 0, "gt_modprep_gq: null pointer or empty");
    modprep_gq_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d, gd, gq, n);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_gq");
    return GT_OK;
}

extern "C" int gt_modprep_style_bwd(const float* g_sn, const float* sn, const float* t, const float* smax, const int* sarg, float* gs, int N, int I, int prenorm,
                                    void* stream) {
    GT_REQUIRE(sn && t && gs, "gt_modprep_style_bwd: null pointer");
    GT_REQUIRE(!prenorm || (smax && sarg), "gt_modprep_style_bwd: pre-normalisation needs smax and sarg");
    modprep_style_bwd_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(g_sn, sn, t, smax, sarg, gs, I, prenorm ? 1 : 0);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_style_bwd");
    return GT_OK;
}

// g_w_dtype: GT_F16 / GT_F32 (ignored when g_w is null)
extern "C" int gt_modprep_weight_bwd(const float* W, const void* g_w, int g_w_dtype, const float* g_wsq, const float* scale, const int* amax, float* gW, int O, int I,
                                     int KK, int prenorm, void* stream) {
    GT_REQUIRE(W && g_wsq && gW, "gt_modprep_weight_bwd: null pointer");
    GT_REQUIRE(!prenorm || (scale && amax), "gt_modprep_weight_bwd: pre-normalisation needs scale and amax");
    GT_REQUIRE(!g_w || g_w_dtype == GT_F16 || g_w_dtype == GT_F32, "gt_modprep_weight_bwd: unsupported gradient dtype code %d", g_w_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    if (g_w && g_w_dtype == GT_F16)
        modprep_weight_bwd_kernel<__half><<<O, 256, 0, st>>>(W, (const __half*)g_w, g_wsq, scale, amax, gW, I, KK, prenorm ? 1 : 0);
    else
        modprep_weight_bwd_kernel<float><<<O, 256, 0, st>>>(W, (const float*)g_w, g_wsq, scale, amax, gW, I, KK, prenorm ? 1 : 0);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_weight_bwd");
    return GT_OK;
}

// Second order of the style side, three stages around the fully-connected products (header of this file).  `sn` is the normalised style
// (the style itself without pre-normalisation); gga may be NULL.
extern "C" int gt_modprep_style_bwd2_a(const float* u, const float* sn, const int* sarg, float* v, float* vs, float* r, int N, int I, int prenorm, void* stream) {
    GT_REQUIRE(u && sn && v && vs && r, "gt_modprep_style_bwd2_a: null pointer");
    GT_REQUIRE(N > 0 && I > 0, "gt_modprep_style_bwd2_a: empty styles");
    GT_REQUIRE(!prenorm || sarg, "gt_modprep_style_bwd2_a: pre-normalisation needs sarg");
    modprep_style_bwd2_a_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(u, sn, sarg, v, vs, r, I, prenorm ? 1 : 0);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_style_bwd2_a");
    return GT_OK;
}

extern "C" int gt_modprep_style_bwd2_b(const float* d, const float* gd, const float* z, const float* smax, float* ggd, float* gq, float* hq, int N, int O,
                                       int prenorm, void* stream) {
    GT_REQUIRE(d && gd && z && ggd && gq && hq, "gt_modprep_style_bwd2_b: null pointer");
    GT_REQUIRE(N > 0 && O > 0, "gt_modprep_style_bwd2_b: empty");
    GT_REQUIRE(!prenorm || smax, "gt_modprep_style_bwd2_b: pre-normalisation needs smax");
    const int total = N * O;
    modprep_style_bwd2_b_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d, gd, z, smax, ggd, gq, hq, total, O, prenorm ? 1 : 0);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_style_bwd2_b");
    return GT_OK;
}

extern "C" int gt_modprep_style_bwd2_c(const float* a, const float* sn, const float* gp, const float* hp, const float* v, const float* r, const float* smax,
                                       const int* sarg, float* gga, float* g2s, float* x1, float* p, int N, int I, int prenorm, void* stream) {
    GT_REQUIRE(sn && gp && hp && v && r && g2s && x1 && p, "gt_modprep_style_bwd2_c: null pointer");
    GT_REQUIRE(N > 0 && I > 0, "gt_modprep_style_bwd2_c: empty styles");
    GT_REQUIRE(!prenorm || (smax && sarg), "gt_modprep_style_bwd2_c: pre-normalisation needs smax and sarg");
    modprep_style_bwd2_c_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(a, sn, gp, hp, v, r, smax, sarg, gga, g2s, x1, p, I, prenorm ? 1 : 0);
    GT_CUDA_LAUNCH_CHECK("gt_modprep_style_bwd2_c");
    return GT_OK;
}
