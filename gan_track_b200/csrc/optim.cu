// optim.cu -- the per-phase parameter update on flat buffers: gradient scrub + Adam as ONE
// kernel over a module's flat fp32 parameter / gradient / moment buffers, and the G_ema update as one kernel.
//
// Reference (S3/training/training_loop_mi_multimodal.py:340-366, S3 = /root/reference/src/models/stylegan3): per phase
//     flat = cat(grads); all_reduce(flat); flat /= num_gpus; nan_to_num(flat, 0, 1e5, -1e5); split back; opt.step()
// and per iteration `p_ema.copy_(p.lerp(p_ema, ema_beta))` for every parameter -- ~400 small tensors, i.e. several hundred
// launches of element-wise kernels.  Here parameters (and their gradients and Adam moments) of a module are views into flat
// buffers; a static table cuts the flat index space into chunks that never straddle a parameter, so that a parameter
// without a gradient in this phase is skipped exactly like torch.optim.Adam skips `grad is None` (no moment decay, no step
// count), and each parameter keeps its own step count for the bias correction.
//
// Arithmetic = torch.optim.Adam (amsgrad off, weight decay 0, maximize off), in its order:
//     g = nan_to_num(grad * grad_scale);  m = m + (g - m)(1 - b1);  v = v b2 + g g (1 - b2)
//     denom = sqrt(v) / sqrt(1 - b2^t) + eps;  p = p - (lr / (1 - b1^t)) m / denom
#include "gt_common.cuh"
#include <math.h>

namespace {

struct Chunk {
    long long pstart;    // element offset into the flat parameter / moment buffers
    long long gstart;    // element offset into the phase's compact gradient buffer
    int count;           // elements
    int seg;             // parameter index
};

__global__ void adam_step_count_kernel(float* __restrict__ steps, const int* __restrict__ active, int nseg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg && active[i]) steps[i] += 1.f;
}

__device__ __forceinline__ float scrub(float g, float scale, float posinf, float neginf) {
    g *= scale;
    if (isnan(g)) return 0.f;
    if (isinf(g)) return g > 0.f ? posinf : neginf;
    return g;
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                                        const float* __restrict__ steps, const int* __restrict__ active, const Chunk* __restrict__ chunks,
                                                        int nchunks, float lr, float b1, float b2, float eps, float grad_scale, float posinf,
                                                        float neginf) {
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const Chunk ch = chunks[c];
        if (!active[ch.seg]) continue;
        const float t = steps[ch.seg];
        const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
        const float step_size = lr / bc1, rsq_bc2 = 1.f / sqrtf(bc2);
        // parameters / moments start on 16-byte boundaries (FlatParams aligns every parameter to 64 floats, chunks are multiples of 4 long
        // except a parameter's last one): 16-byte accesses for p, m, v; the compact gradient buffer has no such alignment -> scalar accesses
        const int n4 = ((ch.pstart & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(m) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(v) & 15) == 0)
                           ? (ch.count >> 2)
                           : 0;
        float4* p4 = reinterpret_cast<float4*>(p + ch.pstart);
        float4* m4 = reinterpret_cast<float4*>(m + ch.pstart);
        float4* v4 = reinterpret_cast<float4*>(v + ch.pstart);
        float* gp = grad + ch.gstart;
        const bool g_aligned = ((ch.gstart & 3) == 0) && ((reinterpret_cast<uintptr_t>(grad) & 15) == 0);
        for (int i = threadIdx.x; i < n4; i += 256) {
            float4 pp = p4[i], mm = m4[i], vv = v4[i];
            float g[4];
            if (g_aligned) {
                const float4 gv = reinterpret_cast<const float4*>(gp)[i];
                g[0] = scrub(gv.x, grad_scale, posinf, neginf);
                g[1] = scrub(gv.y, grad_scale, posinf, neginf);
                g[2] = scrub(gv.z, grad_scale, posinf, neginf);
                g[3] = scrub(gv.w, grad_scale, posinf, neginf);
                reinterpret_cast<float4*>(gp)[i] = make_float4(g[0], g[1], g[2], g[3]);   // the reduced, scrubbed gradient stays readable (tests, stats)
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) g[k] = scrub(gp[4 * i + k], grad_scale, posinf, neginf);
#pragma unroll
                for (int k = 0; k < 4; k++) gp[4 * i + k] = g[k];
            }
            float* pe = reinterpret_cast<float*>(&pp);
            float* me = reinterpret_cast<float*>(&mm);
            float* ve = reinterpret_cast<float*>(&vv);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float m1 = me[k] + (g[k] - me[k]) * (1.f - b1);
                const float v1 = ve[k] * b2 + g[k] * g[k] * (1.f - b2);
                me[k] = m1;
                ve[k] = v1;
                const float denom = sqrtf(v1) * rsq_bc2 + eps;
                pe[k] = pe[k] - step_size * (m1 / denom);
            }
            m4[i] = mm;
            v4[i] = vv;
            p4[i] = pp;
        }
        for (int i = 4 * n4 + threadIdx.x; i < ch.count; i += 256) {
            const long long e = ch.pstart + i, ge = ch.gstart + i;
            const float g = scrub(grad[ge], grad_scale, posinf, neginf);
            grad[ge] = g;
            const float mm = m[e] + (g - m[e]) * (1.f - b1);
            const float vv = v[e] * b2 + g * g * (1.f - b2);
            m[e] = mm;
            v[e] = vv;
            const float denom = sqrtf(vv) * rsq_bc2 + eps;
            p[e] = p[e] - step_size * (mm / denom);
        }
    }
}

__global__ void __launch_bounds__(256) ema_flat_kernel(float* __restrict__ p_ema, const float* __restrict__ p, long long n, float w) {
    // torch._foreach_lerp_(p_ema, p, w): p_ema + w (p - p_ema)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = p_ema[i];
        p_ema[i] = a + w * (p[i] - a);
    }
}

}  // namespace

extern "C" int gt_adam_chunk_bytes(void) { return (int)sizeof(Chunk); }

// chunks: device array of {int64 pstart, int64 gstart, int32 count, int32 seg} built by the caller; steps: float[nseg]; active: int[nseg]
extern "C" int gt_adam_flat(float* p, float* grad, float* m, float* v, float* steps, const int* active, int nseg, const void* chunks, int nchunks, float lr,
                            float beta1, float beta2, float eps, float grad_scale, float posinf, float neginf, void* stream) {
    GT_REQUIRE(p && grad && m && v && steps && active && chunks, "gt_adam_flat: null pointer");
    GT_REQUIRE(nseg > 0 && nchunks > 0, "gt_adam_flat: empty table");
    cudaStream_t st = (cudaStream_t)stream;
    adam_step_count_kernel<<<(nseg + 255) / 256, 256, 0, st>>>(steps, active, nseg);
    GT_CUDA_LAUNCH_CHECK("gt_adam_flat(step count)");
    int grid = gt_num_sms() * 8;
    if (grid > nchunks) grid = nchunks;
    adam_flat_kernel<<<grid, 256, 0, st>>>(p, grad, m, v, steps, active, (const Chunk*)chunks, nchunks, lr, beta1, beta2, eps, grad_scale, posinf, neginf);
    GT_CUDA_LAUNCH_CHECK("gt_adam_flat");
    return GT_OK;
}

extern "C" int gt_ema_flat(float* p_ema, const float* p, long long n, float weight, void* stream) {
    GT_REQUIRE(p_ema && p && n > 0, "gt_ema_flat: bad arguments");
    long long g = (n + 255) / 256;
    if (g > (long long)gt_num_sms() * 16) g = (long long)gt_num_sms() * 16;
    ema_flat_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(p_ema, p, n, weight);
    GT_CUDA_LAUNCH_CHECK("gt_ema_flat");
    return GT_OK;
}
