// augment_params.cu -- the geometric parameter algebra of the ADA pipe in ONE launch.
//
// Gan-track's pipe (S3/training/augment_mi.py:209-312) turns its random draws into a per-sample inverse warp G_inv with ~110 tiny
// tensor ops per call: for every enabled transform a value draw, a gate draw, a compare + where, a stacked 3x3 matrix and a batched
// 3x3 matmul; then the corner / margin computation (:289-299) and five more 3x3 products (:303-312).  The pipe runs three times per
// training iteration, so this is ~330 launches of microsecond kernels.  Here the draws stay with the host framework's generator
// (same calls, same order as the reference -- a fixed seed keeps giving the reference's parameters) and everything downstream of them
// is one single-block kernel: thread b composes G_inv[b] in the reference's order of right-multiplications, the block reduces the
// four margins over the batch (integer ceil), and every thread finishes its 2x3 sampling matrix `theta`.
//
//   G_inv = I @ scale2d_inv(1 - 2 i, 1) @ rotate2d_inv(-pi/2 i90) @ translate2d_inv(round(tx W), round(ty H)) @ scale2d_inv(s, s)
//             @ rotate2d_inv(-th1) @ scale2d_inv(a, 1/a) @ rotate2d_inv(-th2) @ translate2d_inv(fx W, fy H)
#include "gt_common.cuh"

namespace {

struct AugParamsIn {
    // value draw and gate draw of every transform (device pointers; value == nullptr <=> transform disabled)
    const float *xflip_v, *xflip_g;      // rand[B], rand[B]
    const float *rot90_v, *rot90_g;      // rand[B], rand[B]
    const float *xint_v, *xint_g;        // rand[B,2], rand[B,1]
    const float *scale_v, *scale_g;      // randn[B], rand[B]
    const float *rot1_v, *rot1_g;        // rand[B], rand[B]
    const float *aniso_v, *aniso_g;      // randn[B], rand[B]
    const float *rot2_v, *rot2_g;        // rand[B], rand[B]
    const float *xfrac_v, *xfrac_g;      // randn[B,2], rand[B,1]
    const float* p;                      // overall strength (device scalar)
    float m_xflip, m_rot90, m_xint, m_scale, m_rotate, m_aniso, m_xfrac;      // probability multipliers
    float xint_max, scale_std, rotate_max, aniso_std, xfrac_std;
    int B, H, W, hz_pad;
};

struct M3 {
    float a[3][3];
};
__device__ __forceinline__ M3 mul(const M3& x, const M3& y) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.a[i][j] = x.a[i][0] * y.a[0][j] + x.a[i][1] * y.a[1][j] + x.a[i][2] * y.a[2][j];
    return r;
}
__device__ __forceinline__ M3 ident() {
    M3 r = {{{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}}};
    return r;
}
__device__ __forceinline__ M3 scale(float sx, float sy) {
    M3 r = ident();
    r.a[0][0] = sx;
    r.a[1][1] = sy;
    return r;
}
__device__ __forceinline__ M3 translate(float tx, float ty) {
    M3 r = ident();
    r.a[0][2] = tx;
    r.a[1][2] = ty;
    return r;
}
__device__ __forceinline__ M3 rotate(float th) {        // rotate2d(theta) of the pipe: [[cos, sin(-theta), 0], [sin, cos, 0], [0, 0, 1]]
    M3 r = ident();
    const float c = cosf(th), s = sinf(th);
    r.a[0][0] = c;
    r.a[0][1] = sinf(-th);
    r.a[1][0] = s;
    r.a[1][1] = c;
    return r;
}

__global__ void __launch_bounds__(1024) aug_params_kernel(const AugParamsIn in, float* __restrict__ theta, int* __restrict__ margins) {
    __shared__ float red[4][32];
    const int b = threadIdx.x;
    const bool live = b < in.B;
    const float p = *in.p;
    const float PI = 3.14159265358979323846f;
    M3 G = ident();
    if (live) {
        if (in.xflip_v) {
            float i = floorf(in.xflip_v[b] * 2.f);
            i = (in.xflip_g[b] < in.m_xflip * p) ? i : 0.f;
            G = mul(G, scale(1.f / (1.f - 2.f * i), 1.f));
        }
        if (in.rot90_v) {
            float i = floorf(in.rot90_v[b] * 4.f);
            i = (in.rot90_g[b] < in.m_rot90 * p) ? i : 0.f;
            G = mul(G, rotate(-(-PI / 2.f * i)));                      // rotate2d_inv(t) = rotate2d(-t)
        }
        if (in.xint_v) {
            float tx = (in.xint_v[2 * b] * 2.f - 1.f) * in.xint_max, ty = (in.xint_v[2 * b + 1] * 2.f - 1.f) * in.xint_max;
            const bool on = in.xint_g[b] < in.m_xint * p;
            tx = on ? tx : 0.f;
            ty = on ? ty : 0.f;
            G = mul(G, translate(-rintf(tx * (float)in.W), -rintf(ty * (float)in.H)));      // torch.round = round half to even = rintf
        }
        if (in.scale_v) {
            float s = exp2f(in.scale_v[b] * in.scale_std);
            s = (in.scale_g[b] < in.m_scale * p) ? s : 1.f;
            G = mul(G, scale(1.f / s, 1.f / s));
        }
        const float p_rot = 1.f - sqrtf(fminf(fmaxf(1.f - in.m_rotate * p, 0.f), 1.f));
        if (in.rot1_v) {
            float th = (in.rot1_v[b] * 2.f - 1.f) * PI * in.rotate_max;
            th = (in.rot1_g[b] < p_rot) ? th : 0.f;
            G = mul(G, rotate(th));                                     // rotate2d_inv(-th) = rotate2d(th)
        }
        if (in.aniso_v) {
            float s = exp2f(in.aniso_v[b] * in.aniso_std);
            s = (in.aniso_g[b] < in.m_aniso * p) ? s : 1.f;
            G = mul(G, scale(1.f / s, 1.f / (1.f / s)));
        }
        if (in.rot2_v) {
            float th = (in.rot2_v[b] * 2.f - 1.f) * PI * in.rotate_max;
            th = (in.rot2_g[b] < p_rot) ? th : 0.f;
            G = mul(G, rotate(th));
        }
        if (in.xfrac_v) {
            float tx = in.xfrac_v[2 * b] * in.xfrac_std, ty = in.xfrac_v[2 * b + 1] * in.xfrac_std;
            const bool on = in.xfrac_g[b] < in.m_xfrac * p;
            tx = on ? tx : 0.f;
            ty = on ? ty : 0.f;
            G = mul(G, translate(-(tx * (float)in.W), -(ty * (float)in.H)));
        }
    }
    // margins: how far the warped corners reach outside the input, per side, maximised over the batch (:289-299)
    const float cx = (float)(in.W - 1) / 2.f, cy = (float)(in.H - 1) / 2.f;
    float m4[4] = {-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};         // max(-x), max(-y), max(x), max(y)
    if (live) {
        const float xs[4] = {-cx, cx, cx, -cx}, ys[4] = {-cy, -cy, cy, cy};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float x = G.a[0][0] * xs[k] + G.a[0][1] * ys[k] + G.a[0][2], y = G.a[1][0] * xs[k] + G.a[1][1] * ys[k] + G.a[1][2];
            m4[0] = fmaxf(m4[0], -x);
            m4[1] = fmaxf(m4[1], -y);
            m4[2] = fmaxf(m4[2], x);
            m4[3] = fmaxf(m4[3], y);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m4[k] = fmaxf(m4[k], __shfl_xor_sync(0xffffffffu, m4[k], o));
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = m4[k];
    }
    __syncthreads();
    float mf[4];
    const int nw = (blockDim.x + 31) / 32;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float v = red[k][0];
        for (int w = 1; w < nw; w++) v = fmaxf(v, red[k][w]);
        const float c = (k & 1) ? cy : cx;
        v += (float)(in.hz_pad * 2) - c;
        v = fmaxf(v, 0.f);
        v = fminf(v, (float)(((k & 1) ? in.H : in.W) - 1));
        mf[k] = ceilf(v);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) margins[k] = (int)mf[k];
    }
    if (!live) return;
    // :303-312 with the margins kept on the device
    G = mul(translate((mf[0] - mf[2]) / 2.f, (mf[1] - mf[3]) / 2.f), G);
    G = mul(mul(scale(2.f, 2.f), G), scale(1.f / 2.f, 1.f / 2.f));
    G = mul(mul(translate(-0.5f, -0.5f), G), translate(0.5f, 0.5f));
    const float wu = (mf[0] + mf[2] + (float)in.W) * 2.f, hu = (mf[1] + mf[3] + (float)in.H) * 2.f;
    const float ow = (float)((in.W + in.hz_pad * 2) * 2), oh = (float)((in.H + in.hz_pad * 2) * 2);
    G = mul(mul(scale(2.f / wu, 2.f / hu), G), scale(1.f / (2.f / ow), 1.f / (2.f / oh)));
    float* t = theta + (size_t)b * 6;
    t[0] = G.a[0][0];
    t[1] = G.a[0][1];
    t[2] = G.a[0][2];
    t[3] = G.a[1][0];
    t[4] = G.a[1][1];
    t[5] = G.a[1][2];
}

}  // namespace

// ptrs: 16 device pointers in the order of AugParamsIn (value, gate per transform; NULL value = disabled); mult: 7 multipliers;
// ranges: xint_max, scale_std, rotate_max, aniso_std, xfrac_std.  theta: [B,2,3] fp32; margins: int32[4] = x0, y0, x1, y1.
extern "C" int gt_aug_params(const void* const* ptrs, const float* p, const float* mult, const float* ranges, int B, int H, int W, int hz_pad, void* theta,
                             void* margins, void* stream) {
    GT_REQUIRE(ptrs && p && mult && ranges && theta && margins, "gt_aug_params: null pointer");
    GT_REQUIRE(B >= 1 && B <= 1024 && H >= 1 && W >= 1, "gt_aug_params: batch %d must be in [1, 1024]", B);
    AugParamsIn in;
    const float** f = (const float**)&in;
    for (int i = 0; i < 16; i++) f[i] = (const float*)ptrs[i];
    for (int i = 0; i < 8; i++) GT_REQUIRE(f[2 * i] == nullptr || f[2 * i + 1] != nullptr, "gt_aug_params: transform %d has a value draw but no gate draw", i);
    in.p = p;
    in.m_xflip = mult[0];
    in.m_rot90 = mult[1];
    in.m_xint = mult[2];
    in.m_scale = mult[3];
    in.m_rotate = mult[4];
    in.m_aniso = mult[5];
    in.m_xfrac = mult[6];
    in.xint_max = ranges[0];
    in.scale_std = ranges[1];
    in.rotate_max = ranges[2];
    in.aniso_std = ranges[3];
    in.xfrac_std = ranges[4];
    in.B = B;
    in.H = H;
    in.W = W;
    in.hz_pad = hz_pad;
    const int threads = ((B + 31) / 32) * 32;
    aug_params_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(in, (float*)theta, (int*)margins);
    GT_CUDA_LAUNCH_CHECK("gt_aug_params");
    return GT_OK;
}
