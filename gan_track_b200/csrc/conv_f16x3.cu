// conv_f16x3.cu -- operand preparation and final reduction of the fp32 convolutions on the fp16 tensor-core path ("fp16 x 3").
//
// The reference keeps the low-resolution blocks in true fp32 (S3/training/networks_stylegan2.py:486, 756; TF32 is switched off,
// S3/training/training_loop_mi_multimodal.py:169-170) and hands their convolutions to the library (OPS/conv2d_gradfix.py:37-45).
// Here they run on the same tcgen05 implicit-GEMM kernels as the fp16 layers:
//
//   * every fp32 tensor v is scaled by a power of two s = 2^(13 - floor(log2 max|v|)) (so max|v| s is in [2^13, 2^14), exact) and split
//     into hi = fp16(v s) and lo = fp16(v s - hi): hi + lo carries 22 significant bits of v s (elements far below the tensor maximum
//     lose relative but not absolute precision -- the error stays ~2^-36 of the tensor maximum);
//   * x w  ~=  (x_hi w_lo + x_lo w_hi + x_hi w_hi) / (s_x s_w); the dropped x_lo w_lo term is ~2^-22 relative.  The three products
//     are ONE fp16 convolution over a 3x wider channel axis: activations [N,H,W,3C] = [hi | lo | hi], weights
//     [tap][Cout][3C] = [lo | hi | hi] (fp16 x fp16 products are exact in the fp32 accumulator);
//   * the tensor core adds every MMA into its fp32 accumulator with truncation, an error proportional to the accumulator's magnitude
//     and to the number of updates (measured: 7e-6 at ~450 updates, 9e-6 at 288 with the main term first).  Two measures keep it
//     near 2e-6: the channel chunks are ordered cross terms first, main term last (the kernels walk channel chunks in the OUTER loop,
//     taps inside), so that the 2/3 of the updates that carry the 2^-11-times smaller cross terms happen while the accumulator is
//     still small; and the reduction is split by kernel row (K-split phases of gt_conv2d_igemm_f16_f32out), the partial sums being
//     added here in fp32 with round-to-nearest together with the 1 / (s_x s_w) rescale;
//   * weight gradients use the same split along the BATCH axis: dW = sum_px U S with U' = (u_hi, u_lo, u_hi), S' = (s_lo, s_hi, s_hi)
//     as 3N images (cross terms first again: every split-K slice walks the images in increasing order), and the wgrad kernels are
//     asked for enough split-K slices that no accumulator sees more than ~256 main-term updates.
//
// The maximum is taken by an integer atomicMax on the bit pattern of |v| (order independent -> deterministic).
#include "gt_common.cuh"

namespace {

__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long s0, long long s1, long long s2, long long s3, int d1, int d2, int d3,
                                                   long long total, uint32_t* __restrict__ out) {
    uint32_t m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int i3 = (int)(i % d3);
        long long r = i / d3;
        const int i2 = (int)(r % d2);
        r /= d2;
        const int i1 = (int)(r % d1);
        const long long i0 = r / d1;
        const uint32_t b = __float_as_uint(x[i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3]) & 0x7fffffffu;
        m = b > m ? b : m;              // NaNs compare high and end up as scale 1 (they poison the result either way)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t v = __shfl_xor_sync(0xffffffffu, m, o);
        m = v > m ? v : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// dense tensors (any permutation of a contiguous block): the maximum does not care about the order -> flat 16-byte loads
__global__ void __launch_bounds__(256) amax_dense_kernel(const float* __restrict__ x, long long n, uint32_t* __restrict__ out) {
    uint32_t m = 0;
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x4[i];
        const uint32_t a = __float_as_uint(v.x) & 0x7fffffffu, b = __float_as_uint(v.y) & 0x7fffffffu, c = __float_as_uint(v.z) & 0x7fffffffu,
                       d = __float_as_uint(v.w) & 0x7fffffffu;
        m = max(max(m, max(a, b)), max(c, d));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = max(m, __float_as_uint(x[(n4 << 2) + threadIdx.x]) & 0x7fffffffu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ uint32_t sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; w++) m = max(m, sm[w]);
        if (m) atomicMax(out, m);
    }
}

// fast path of split_act_kernel: x dense channels-last (s_c == 1), C == Cp a multiple of 8: one thread = 8 channels of one pixel
__global__ void __launch_bounds__(256) split_act_nhwc8_kernel(const float* __restrict__ x, uint32_t npix, uint32_t c8n, uint32_t pix_per_img, int N, int layout,
                                                              const uint32_t* __restrict__ amax, __half* __restrict__ out) {
    const float sc = gt_scale_from_amax_bits(*amax);
    const uint32_t total = npix * c8n;
    const uint32_t Cp = c8n * 8;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t pix = i / c8n, c = (i - pix * c8n) * 8;
        const float4 a = *reinterpret_cast<const float4*>(x + (size_t)pix * Cp + c), b = *reinterpret_cast<const float4*>(x + (size_t)pix * Cp + c + 4);
        const float v[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, b.x * sc, b.y * sc, b.z * sc, b.w * sc};
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            hi[k] = __float2half_rn(v[k]);
            lo[k] = __float2half_rn(v[k] - __half2float(hi[k]));
        }
        const uint4 H = *reinterpret_cast<const uint4*>(hi), L = *reinterpret_cast<const uint4*>(lo);
        if (layout == 0) {
            __half* o = out + (size_t)pix * (3 * Cp) + c;
            *reinterpret_cast<uint4*>(o) = H;
            *reinterpret_cast<uint4*>(o + Cp) = L;
            *reinterpret_cast<uint4*>(o + 2 * Cp) = H;
        } else {
            __half* o = out + (size_t)pix * Cp + c;
            const size_t plane = (size_t)N * pix_per_img * Cp;
            *reinterpret_cast<uint4*>(o) = layout == 1 ? H : L;
            *reinterpret_cast<uint4*>(o + plane) = layout == 1 ? L : H;
            *reinterpret_cast<uint4*>(o + 2 * plane) = H;
        }
    }
}

// fast path of pack_weight_kernel for kernels of up to 9 taps: one thread = one (co, ci) pair, all taps (its taps are KH*KW consecutive
// floats when the weight is contiguous)
__global__ void __launch_bounds__(256) pack_weight_pair_kernel(const float* __restrict__ w, long long s_co, long long s_ci, long long s_r, long long s_s, int Cout,
                                                               int Cin, int KH, int KW, int Coutp, int Cinp, const uint32_t* __restrict__ amax,
                                                               __half* __restrict__ out) {
    const float sc = gt_scale_from_amax_bits(*amax);
    const uint32_t total = (uint32_t)Coutp * (uint32_t)Cinp;
    const __half z = __float2half_rn(0.f);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t co = i / (uint32_t)Cinp, ci = i - co * (uint32_t)Cinp;
        const bool live = ci < (uint32_t)Cin && co < (uint32_t)Cout;
        const float* wp = w + (long long)co * s_co + (long long)ci * s_ci;
        for (int r = 0; r < KH; r++)
            for (int q = 0; q < KW; q++) {
                __half hi = z, lo = z;
                if (live) {
                    const float v = wp[r * s_r + q * s_s] * sc;
                    hi = __float2half_rn(v);
                    lo = __float2half_rn(v - __half2float(hi));
                }
                __half* o = out + ((size_t)(r * KW + q) * Coutp + co) * (3ull * Cinp) + ci;
                o[0] = lo;
                o[Cinp] = hi;
                o[2 * Cinp] = hi;
            }
    }
}

// activations: x [N,C,H,W] (element strides) -> out fp16.
//   layout 0 (channel concat, operand of the forward / data-gradient kernels): out[n][h][w][3*Cp] = [hi | lo | hi]
//   layout 1 (batch concat, U operand of the weight-gradient kernels):         out[3N][h][w][Cp], images (hi, lo, hi)
//   layout 2 (batch concat, S operand):                                        out[3N][h][w][Cp], images (lo, hi, hi)
// Cp >= C is the channel count padded to a multiple of 64 (padding written as zero).
__global__ void __launch_bounds__(256) split_act_kernel(const float* __restrict__ x, long long s_n, long long s_c, long long s_h, long long s_w, int N, int C, int H,
                                                        int W, int Cp, int layout, const uint32_t* __restrict__ amax, __half* __restrict__ out) {
    const float sc = gt_scale_from_amax_bits(*amax);
    const long long total = (long long)N * H * W * Cp;
    const long long img = (long long)H * W * Cp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        long long r = i / Cp;
        const int w = (int)(r % W);
        r /= W;
        const int h = (int)(r % H);
        const int n = (int)(r / H);
        __half hi = __float2half_rn(0.f), lo = hi;
        if (c < C) {
            const float v = x[n * s_n + c * s_c + h * s_h + w * s_w] * sc;
            hi = __float2half_rn(v);
            lo = __float2half_rn(v - __half2float(hi));
        }
        if (layout == 0) {
            __half* o = out + (i / Cp) * (3ll * Cp) + c;
            o[0] = hi;
            o[Cp] = lo;
            o[2 * Cp] = hi;
        } else {
            const long long pix = i - (long long)n * img;          // offset inside the image
            __half* o = out + (long long)n * img + pix;
            const long long plane = (long long)N * img;
            o[0] = layout == 1 ? hi : lo;
            o[plane] = layout == 1 ? lo : hi;
            o[2 * plane] = hi;
        }
    }
}

// weights: w[co * s_co + ci * s_ci + r * s_r + s * s_s] -> out[t][Coutp][3*Cinp] = [lo | hi | hi], zero padded: pack_weight_pair_kernel above

// y[i] = (sum_s slab[s][i]) / (scale_a * scale_b), i over a dense [rows][Cp] fp32 array of which the first C columns are kept:
// y is [rows][C] dense (the NHWC output of the convolution)
__global__ void __launch_bounds__(256) slab_reduce_kernel(const float* __restrict__ slabs, long long slab_stride, int nslabs, long long rows, int Cp, int C,
                                                          const uint32_t* __restrict__ amax_a, const uint32_t* __restrict__ amax_b, float* __restrict__ y) {
    const float inv = (1.f / gt_scale_from_amax_bits(*amax_a)) * (1.f / gt_scale_from_amax_bits(*amax_b));      // powers of two: exact
    const int c4 = (C % 4 == 0) ? C / 4 : 0;          // 16-byte stores need every output row 16-byte aligned
    const long long total = rows * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / c4;
        const int c = (int)(i - row * c4) * 4;
        const float* p = slabs + row * Cp + c;
        float4 acc = *reinterpret_cast<const float4*>(p);
        for (int s = 1; s < nslabs; s++) {
            const float4 v = *reinterpret_cast<const float4*>(p + s * slab_stride);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
        acc.x *= inv;
        acc.y *= inv;
        acc.z *= inv;
        acc.w *= inv;
        *reinterpret_cast<float4*>(y + row * C + c) = acc;
    }
    // channel counts that are not a multiple of 4 (the 513-channel data gradient below the minibatch-std layer): scalar path
    const int tail = C - c4 * 4;
    if (tail) {
        const long long tt = rows * tail;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tt; i += (long long)gridDim.x * blockDim.x) {
            const long long row = i / tail;
            const int c = c4 * 4 + (int)(i - row * tail);
            float acc = 0.f;
            for (int s = 0; s < nslabs; s++) acc += slabs[s * slab_stride + row * Cp + c];
            y[row * C + c] = acc * inv;
        }
    }
}

int grid_for(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)gt_num_sms() * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace

// max |x| over a tensor of up to four dimensions (element strides) as the bit pattern of the fp32 value; `amax_bits` (one uint32 in device
// memory) is cleared on the stream first
extern "C" int gt_f16x3_amax(const void* x, long long s0, long long s1, long long s2, long long s3, int d0, int d1, int d2, int d3, void* amax_bits,
                             void* stream) {
    GT_REQUIRE(x && amax_bits, "gt_f16x3_amax: null pointer");
    GT_REQUIRE(d0 > 0 && d1 > 0 && d2 > 0 && d3 > 0, "gt_f16x3_amax: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(amax_bits, 0, 4, st) != cudaSuccess) {
        gt_set_error("gt_f16x3_amax: memset failed");
        return GT_ERR_CUDA;
    }
    const long long total = (long long)d0 * d1 * d2 * d3;
    // dense in some dimension order (strides are a permutation of a contiguous layout) and 16-byte aligned: flat vector loads
    bool dense = ((uintptr_t)x & 15) == 0;
    {
        long long st_[4] = {s0, s1, s2, s3};
        int dm[4] = {d0, d1, d2, d3};
        long long expect = 1;
        bool used[4] = {false, false, false, false};
        for (int k = 0; k < 4 && dense; k++) {
            int pick = -1;
            for (int j = 0; j < 4; j++)
                if (!used[j] && (dm[j] == 1 || st_[j] == expect)) {
                    if (pick < 0 || dm[j] == 1) pick = j;
                }
            if (pick < 0) dense = false;
            else {
                used[pick] = true;
                expect *= dm[pick];
            }
        }
    }
    if (dense) {
        long long g = (total / 4 + 255) / 256;
        const long long cap = (long long)gt_num_sms() * 8;
        if (g > cap) g = cap;
        if (g < 1) g = 1;
        amax_dense_kernel<<<(int)g, 256, 0, st>>>((const float*)x, total, (uint32_t*)amax_bits);
    } else
        amax_kernel<<<grid_for(total), 256, 0, st>>>((const float*)x, s0, s1, s2, s3, d1, d2, d3, total, (uint32_t*)amax_bits);
    GT_CUDA_LAUNCH_CHECK("gt_f16x3_amax");
    return GT_OK;
}

extern "C" int gt_f16x3_split_act(const void* x, long long s_n, long long s_c, long long s_h, long long s_w, int N, int C, int H, int W, int Cp, int layout,
                                  const void* amax_bits, void* out, void* stream) {
    GT_REQUIRE(x && out && amax_bits, "gt_f16x3_split_act: null pointer");
    GT_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && Cp >= C && Cp % 64 == 0, "gt_f16x3_split_act: bad shape (C %d, padded %d)", C, Cp);
    GT_REQUIRE(layout >= 0 && layout <= 2, "gt_f16x3_split_act: layout must be 0 (channel concat), 1 (batch concat, U role) or 2 (batch concat, S role)");
    const long long total = (long long)N * H * W * Cp;
    if (C == Cp && s_c == 1 && s_w == C && s_h == (long long)W * C && s_n == (long long)H * W * C && ((uintptr_t)x & 15) == 0 && total < (1ll << 31)) {
        const uint32_t npix = (uint32_t)(N * H * W), c8n = (uint32_t)(Cp / 8);
        split_act_nhwc8_kernel<<<grid_for(total / 8), 256, 0, (cudaStream_t)stream>>>((const float*)x, npix, c8n, (uint32_t)(H * W), N, layout,
                                                                                    (const uint32_t*)amax_bits, (__half*)out);
        GT_CUDA_LAUNCH_CHECK("gt_f16x3_split_act");
        return GT_OK;
    }
    split_act_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const float*)x, s_n, s_c, s_h, s_w, N, C, H, W, Cp, layout, (const uint32_t*)amax_bits,
                                                                        (__half*)out);
    GT_CUDA_LAUNCH_CHECK("gt_f16x3_split_act");
    return GT_OK;
}

extern "C" int gt_f16x3_pack_weight(const void* w, long long s_co, long long s_ci, long long s_r, long long s_s, int Cout, int Cin, int KH, int KW, int Coutp,
                                    int Cinp, const void* amax_bits, void* out, void* stream) {
    GT_REQUIRE(w && out && amax_bits, "gt_f16x3_pack_weight: null pointer");
    GT_REQUIRE(Cout > 0 && Cin > 0 && KH > 0 && KW > 0 && Coutp >= Cout && Cinp >= Cin && Coutp % 64 == 0 && Cinp % 64 == 0, "gt_f16x3_pack_weight: bad shape");
    pack_weight_pair_kernel<<<grid_for((long long)Coutp * Cinp), 256, 0, (cudaStream_t)stream>>>((const float*)w, s_co, s_ci, s_r, s_s, Cout, Cin, KH, KW, Coutp,
                                                                                                 Cinp, (const uint32_t*)amax_bits, (__half*)out);
    GT_CUDA_LAUNCH_CHECK("gt_f16x3_pack_weight");
    return GT_OK;
}

extern "C" int gt_f16x3_slab_reduce(const void* slabs, long long slab_stride, int nslabs, long long rows, int Cp, int C, const void* amax_a,
                                    const void* amax_b, void* y, void* stream) {
    GT_REQUIRE(slabs && y && amax_a && amax_b, "gt_f16x3_slab_reduce: null pointer");
    GT_REQUIRE(nslabs >= 1 && rows > 0 && C > 0 && Cp >= C && Cp % 4 == 0 && slab_stride % 4 == 0, "gt_f16x3_slab_reduce: bad shape");
    slab_reduce_kernel<<<grid_for(rows * ((C + 3) / 4)), 256, 0, (cudaStream_t)stream>>>((const float*)slabs, slab_stride, nslabs, rows, Cp, C,
                                                                                        (const uint32_t*)amax_a, (const uint32_t*)amax_b, (float*)y);
    GT_CUDA_LAUNCH_CHECK("gt_f16x3_slab_reduce");
    return GT_OK;
}
