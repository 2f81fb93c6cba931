// hot_act.cuh -- lean arithmetic of the two activations on the StyleGAN2 path (linear, lrelu) for the HBM-streaming
// kernels.  Same formulas and the same fp32 operation order as the reference kernel (OPS/bias_act.cu:60-142; OPS =
// /root/reference/src/models/stylegan3/torch_utils/ops), evaluated two lanes at a time with the sm_100 packed fp32
// instructions (fmul2 / fadd2 are IEEE per lane, so results are bit-identical to the scalar form).  The streaming
// kernels are issue-bound before they are HBM-bound (ncu: 150 warp instructions per 16-byte vector with the generic
// nine-activation evaluator), hence: compile-time GRAD and CLAMP, max() instead of compare+select where exact,
// NaN-propagating min/max for the clamp, packed half2 <-> float2 conversions.
#pragma once
#include "gt_common.cuh"

namespace hot {

enum { LINEAR = 1, LRELU = 3 };

__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

struct Params {
    float alpha, gain, clampv;
    bool alpha01;        // 0 <= alpha <= 1: max(v, alpha v) == (v > 0 ? v : alpha v), exactly, incl. signed zeros and NaN
    uint32_t sgn;        // sign bit of gain: (yref / gain > 0) == ((yref ^ sgn) > 0) for gain != 0
};
__host__ __device__ inline Params make_params(float alpha, float gain, float clampv) {
    Params p;
    p.alpha = alpha;
    p.gain = gain;
    p.clampv = clampv;
    p.alpha01 = alpha >= 0.f && alpha <= 1.f;
    p.sgn = gain < 0.f ? 0x80000000u : 0u;
    return p;
}

// forward: clamp(act(v) * gain), v = x + b
template <int ACT, bool CLAMP>
__device__ __forceinline__ float2 fwd(float2 v, const Params& p) {
    float2 t = v;
    if (ACT == LRELU) {
        const float2 va = __fmul2_rn(v, make_float2(p.alpha, p.alpha));
        if (p.alpha01) {
            t.x = max_nan(v.x, va.x);
            t.y = max_nan(v.y, va.y);
        } else {
            t.x = v.x > 0.f ? v.x : va.x;
            t.y = v.y > 0.f ? v.y : va.y;
        }
    }
    float2 r = __fmul2_rn(t, make_float2(p.gain, p.gain));
    if (CLAMP) {
        r.x = min_nan(max_nan(r.x, -p.clampv), p.clampv);
        r.y = min_nan(max_nan(r.y, -p.clampv), p.clampv);
    }
    return r;
}

// first derivative pass: dy * act'(.) * gain, zero where the forward output sat on the clamp; decided from the saved
// forward output yr like the reference (yy = yr / gain; yy > 0)
template <int ACT, bool CLAMP>
__device__ __forceinline__ float2 bwd(float2 g, float2 yr, const Params& p) {
    float2 t = g;
    if (ACT == LRELU) {
        const float2 ga = __fmul2_rn(g, make_float2(p.alpha, p.alpha));
        t.x = (__uint_as_float(__float_as_uint(yr.x) ^ p.sgn) > 0.f) ? g.x : ga.x;
        t.y = (__uint_as_float(__float_as_uint(yr.y) ^ p.sgn) > 0.f) ? g.y : ga.y;
    }
    float2 r = __fmul2_rn(t, make_float2(p.gain, p.gain));
    if (CLAMP) {
        r.x = (yr.x > -p.clampv && yr.x < p.clampv) ? r.x : 0.f;
        r.y = (yr.y > -p.clampv && yr.y < p.clampv) ? r.y : 0.f;
    }
    return r;
}

// ---- 16-byte vector <-> float2 lanes ------------------------------------------------------------------------------
template <class T> struct Lanes;   // NP float2 pairs per 16-byte vector
template <> struct Lanes<__half> {
    static constexpr int NP = 4;
    static __device__ __forceinline__ float2 get(const Vec16<__half>& v, int i) { return __half22float2(reinterpret_cast<const __half2*>(v.v)[i]); }
    static __device__ __forceinline__ void set(Vec16<__half>& v, int i, float2 f) { reinterpret_cast<__half2*>(v.v)[i] = __float22half2_rn(f); }
    // the value as stored (after rounding to the I/O type)
    static __device__ __forceinline__ float2 stored(const Vec16<__half>& v, int i) { return get(v, i); }
};
template <> struct Lanes<float> {
    static constexpr int NP = 2;
    static __device__ __forceinline__ float2 get(const Vec16<float>& v, int i) { return reinterpret_cast<const float2*>(v.v)[i]; }
    static __device__ __forceinline__ void set(Vec16<float>& v, int i, float2 f) { reinterpret_cast<float2*>(v.v)[i] = f; }
    static __device__ __forceinline__ float2 stored(const Vec16<float>& v, int i) { return get(v, i); }
};

}  // namespace hot
