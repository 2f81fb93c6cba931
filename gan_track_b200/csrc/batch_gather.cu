// batch_gather.cu -- a training batch from a device-resident packed shard in one launch.
//
// Replaces, per iteration, the reference's DataLoader path: unzip + unpickle of every slice on host workers
// (S3/training/dataset_mi_multimodal.py:255-268; S3 = /root/reference/src/models/stylegan3), the x-flip copy (:113-116), batch
// collation, the pageable host->device copy and `real_img.to(float32) / 127.5 - 1`
// (S3/training/training_loop_mi_multimodal.py:313-317).
//     out[b,c,y,x] = decode(src[raw_idx[idx[b]], c, y, xflip[idx[b]] ? W-1-x : x]) / scale + shift
// decode: float32 / float16 as is; uint16 = stored round(v * 257) -> v = u / 257.  True divisions (not reciprocal multiplies) so
// that float32 shards reproduce the reference's arithmetic bit for bit.  An index outside [0, n_idx) yields NaN pixels.
#include "gt_common.cuh"

#include <cuda_fp16.h>

namespace {

enum { BG_F32 = 0, BG_F16 = 1, BG_U16 = 3 };

template <int DT>
__device__ __forceinline__ float bg_load(const void* src, long long i) {
    if (DT == BG_F32) return ((const float*)src)[i];
    if (DT == BG_F16) return __half2float(((const __half*)src)[i]);
    return __fdiv_rn((float)((const unsigned short*)src)[i], 257.f);
}

template <int DT>
__global__ void __launch_bounds__(256) batch_gather_kernel(const void* __restrict__ src, const long long* __restrict__ idx, const long long* __restrict__ raw_idx,
                                                           const unsigned char* __restrict__ xflip, float* __restrict__ dst, int C, int H, int W, int n_src,
                                                           int n_idx, float scale, float shift) {
    const int b = blockIdx.y;
    const long long id = idx[b];
    const long long plane = (long long)C * H * W;
    float* out = dst + (long long)b * plane;
    const bool ok = id >= 0 && id < n_idx;
    const long long r = ok ? raw_idx[id] : 0;
    const bool valid = ok && r >= 0 && r < n_src;
    const bool flip = valid && xflip[id] != 0;
    const long long base = r * plane;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < plane; e += (long long)gridDim.x * 256) {
        float v = __int_as_float(0x7fc00000);
        if (valid) {
            long long s = e;
            if (flip) {
                const long long row = e / W;
                const int x = (int)(e - row * W);
                s = row * W + (W - 1 - x);
            }
            v = __fadd_rn(__fdiv_rn(bg_load<DT>(src, base + s), scale), shift);
        }
        out[e] = v;
    }
}

}  // namespace

// src_dtype: 0 float32, 1 float16, 3 uint16 (see above).  idx [B] int64 dataset indices; raw_idx [n_idx] int64 and xflip [n_idx] uint8 are
// the dataset's index tables (max_size subset and flip doubling).  dst [B,C,H,W] float32.
extern "C" int gt_batch_gather(const void* src, int src_dtype, const long long* idx, const long long* raw_idx, const unsigned char* xflip, float* dst, int B, int C,
                               int H, int W, int n_src, int n_idx, float scale, float shift, void* stream) {
    GT_REQUIRE(src && idx && raw_idx && xflip && dst, "gt_batch_gather: null pointer");
    GT_REQUIRE(B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0 && n_src > 0 && n_idx > 0, "gt_batch_gather: empty batch or shard");
    GT_REQUIRE(scale != 0.f, "gt_batch_gather: scale must be non-zero");
    const long long plane = (long long)C * H * W;
    long long bx = (plane + 255) / 256;
    const long long cap = ((long long)gt_num_sms() * 8 + B - 1) / B;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    switch (src_dtype) {
        case BG_F32: batch_gather_kernel<BG_F32><<<grid, 256, 0, st>>>(src, idx, raw_idx, xflip, dst, C, H, W, n_src, n_idx, scale, shift); break;
        case BG_F16: batch_gather_kernel<BG_F16><<<grid, 256, 0, st>>>(src, idx, raw_idx, xflip, dst, C, H, W, n_src, n_idx, scale, shift); break;
        case BG_U16: batch_gather_kernel<BG_U16><<<grid, 256, 0, st>>>(src, idx, raw_idx, xflip, dst, C, H, W, n_src, n_idx, scale, shift); break;
        default: gt_set_error("gt_batch_gather: unsupported source dtype code %d", src_dtype); return GT_ERR_ARG;
    }
    GT_CUDA_LAUNCH_CHECK("gt_batch_gather");
    return GT_OK;
}
