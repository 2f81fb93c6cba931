// fc.cu -- the fully-connected layers of the path (mapping networks, style affines, discriminator epilogue) at training
// batch sizes: fp32, M = batch <= 64 rows.
//
// Reference: FullyConnectedLayer.forward (S3/training/networks_stylegan2.py:115-126; S3 = /root/reference/src/models/stylegan3)
//     w = weight * weight_gain;  b = bias * bias_gain;  y = addmm(b, x, w.t())        (+ bias_act for non-linear layers)
// i.e. per layer and pass two scaling kernels, a library GEMM whose tiles are sized for M >= 128 (the 512x512 layers run
// 11 us, the 8192 -> 512 discriminator layer 110-180 us per GEMM in the step profile, profiles/r01b_step_profile.txt) and
// often a split-K reduction.  With M <= 64 the problem is a stream over the weight matrix; the three kernels below read
// (or write) it exactly once, coalesced, and fold the gains in:
//     gt_fc_fwd     y[m,o]  = wgain * sum_i x[m,i]  w[o,i] + bgain * b[o]
//     gt_fc_dgrad   dx[m,i] = wgain * sum_o dy[m,o] w[o,i]
//     gt_fc_wgrad   dw[o,i] = wgain * sum_m dy[m,o] x[m,i];   db[o] = bgain * sum_m dy[m,o]
// The family is closed under differentiation (the derivative of each is another of the three), which is what the
// path-length and R1 double backwards need.  All reductions have a fixed order.
#include "gt_common.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int FC_KC = 128;     // k-chunk of the forward kernel (one float4 per lane)

// sum over the 32 lanes of MB per-lane values; afterwards acc[0] on lane l is the total of value l (+32 for the upper half)
template <int N>
__device__ __forceinline__ void warp_transpose_sum(float (&acc)[N], int lane) {
#pragma unroll
    for (int off = N / 2; off >= 1; off >>= 1) {
#pragma unroll
        for (int j = 0; j < off; j++) {
            const bool up = (lane & off) != 0;
            const float send = up ? acc[j] : acc[j + off];
            const float keep = up ? acc[j + off] : acc[j];
            acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// One warp per output feature, 8 features per CTA; x staged per 128-wide k-chunk.  The reduction dimension is split over
// the CTAs of a thread-block cluster (blockIdx.y = cluster rank = k-slice): a 512x512 layer would otherwise occupy 64 CTAs
// that each walk four dependent load -> stage -> multiply rounds (8.9 us, latency-bound); with a 4-CTA cluster every CTA
// does one round and the partial sums meet in distributed shared memory, added in rank order (fixed summation order).
template <int MB>
__global__ void __launch_bounds__(256) fc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                     float* __restrict__ y, int M, int I, int O, float wgain, float bgain, int k_per_rank) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), nrank = (int)cluster.num_blocks();
    __shared__ float4 xs[MB][FC_KC / 4];
    __shared__ float part[8][MB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 8 + warp;
    const int kbeg = rank * k_per_rank, kend = min(I, kbeg + k_per_rank);
    float acc[MB];
#pragma unroll
    for (int m = 0; m < MB; m++) acc[m] = 0.f;
    // software pipeline: the x chunk and the weight vector of round r+1 are in flight (registers) while round r is consumed
    constexpr int XPT = MB * (FC_KC / 4) / 256;        // float4 of x per thread per chunk
    float4 xpre[XPT];
    float4 wpre;
    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < XPT; j++) {
            const int idx = threadIdx.x + j * 256;
            const int m = idx / (FC_KC / 4), q = idx - m * (FC_KC / 4);
            const int kk = k0 + q * 4;
            xpre[j] = (m < M && kk < kend) ? *reinterpret_cast<const float4*>(x + (long long)m * I + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int k = k0 + lane * 4;
        wpre = (o < O && k < kend) ? *reinterpret_cast<const float4*>(w + (long long)o * I + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += FC_KC) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < XPT; j++) {
            const int idx = threadIdx.x + j * 256;
            xs[idx / (FC_KC / 4)][idx % (FC_KC / 4)] = xpre[j];
        }
        const float4 wv = wpre;
        __syncthreads();
        if (k0 + FC_KC < kend) fetch(k0 + FC_KC);
#pragma unroll
        for (int m = 0; m < MB; m++) {
            const float4 xv = xs[m][lane];
            acc[m] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
        }
    }
    if (MB >= 32) {
        float lo[32];
#pragma unroll
        for (int m = 0; m < 32; m++) lo[m] = acc[m];
        warp_transpose_sum<32>(lo, lane);
        part[warp][lane] = lo[0];
        if (MB == 64) {
            float hi[32];
#pragma unroll
            for (int m = 0; m < 32; m++) hi[m] = acc[(MB == 64 ? 32 : 0) + m];
            warp_transpose_sum<32>(hi, lane);
            part[warp][(MB == 64 ? 32 : 0) + lane] = hi[0];
        }
    } else {
        // MB in {8, 16}: replicate to 32 slots would waste shuffles; plain butterfly per value
#pragma unroll
        for (int m = 0; m < MB; m++) {
            const float s = warp_sum(acc[m]);
            if (lane == 0) part[warp][m] = s;
        }
    }
    cluster.sync();
    // rank r finishes the output features ow with ow % nrank == r of this CTA column
    for (int idx = threadIdx.x; idx < 8 * MB; idx += 256) {
        const int ow = idx / MB, m = idx - ow * MB;
        const int oo = blockIdx.x * 8 + ow;
        if ((ow % nrank) == rank && oo < O && m < M) {
            float s = 0.f;
            for (int r = 0; r < nrank; r++) s += cluster.map_shared_rank(&part[0][0], r)[ow * MB + m];
            y[(long long)m * O + oo] = s * wgain + (b ? b[oo] * bgain : 0.f);
        }
    }
    cluster.sync();        // keep this CTA's partial sums alive until every rank has read them
}

// Lanes walk 32 input features (coalesced rows of w), the 8 warps split the output features of a round; the output
// features are split over the CTAs of a cluster (blockIdx.y = rank), partial sums combined through distributed shared
// memory in rank order.  Without the split a 512x512 layer is 16 CTAs walking 512 rows each: 16 us.
template <int MB>
__global__ void __launch_bounds__(256) fc_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int M, int I,
                                                       int O, float wgain, int o_per_rank) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), nrank = (int)cluster.num_blocks();
    constexpr int OC = 64;                       // output features staged per round (8 weight loads in flight per thread)
    __shared__ __align__(16) float dys[OC][MB];
    __shared__ float red[8][8][33];
    __shared__ float part[MB][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 32 + lane;
    const int obeg = rank * o_per_rank, oend = min(O, obeg + o_per_rank);
    float acc[MB];
#pragma unroll
    for (int m = 0; m < MB; m++) acc[m] = 0.f;
    for (int o0 = obeg; o0 < oend; o0 += OC) {
        // this warp's 8 weight rows of the round are requested together (8 loads in flight) before the dy staging barrier
        float wv[OC / 8];
#pragma unroll
        for (int j = 0; j < OC / 8; j++) {
            const int o = o0 + warp + 8 * j;
            wv[j] = (o < oend && i < I) ? __ldg(w + (long long)o * I + i) : 0.f;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < OC * MB; idx += 256) {
            const int m = idx / OC, oo = idx - m * OC;         // consecutive threads -> consecutive o: coalesced reads of dy[m, :]
            // rows of dys are MB floats apart, so a plain [oo][m] store would put all 32 lanes on one bank; XOR the 4-float
            // group index with the row (groups of 4 m stay contiguous for the 16-byte reads below)
            dys[oo][m ^ (((oo & (MB / 4 - 1))) << 2)] = (m < M && o0 + oo < oend) ? dy[(long long)m * O + o0 + oo] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < OC / 8; j++) {
            const int oo = warp + 8 * j;
            const int sw = (oo & (MB / 4 - 1)) << 2;
#pragma unroll
            for (int m4 = 0; m4 < MB; m4 += 4) {
                const float4 d4 = *reinterpret_cast<const float4*>(&dys[oo][m4 ^ sw]);
                acc[m4 + 0] += d4.x * wv[j];
                acc[m4 + 1] += d4.y * wv[j];
                acc[m4 + 2] += d4.z * wv[j];
                acc[m4 + 3] += d4.w * wv[j];
            }
        }
    }
    // combine the 8 warps in warp order, 8 rows at a time (keeps the scratch at 8.4 KB for any MB)
#pragma unroll
    for (int mb = 0; mb < MB; mb += 8) {
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < 8; mm++) red[warp][mm][lane] = acc[mb + mm];
        __syncthreads();
        {
            const int mm = threadIdx.x >> 5, l = threadIdx.x & 31;       // 256 threads = 8 rows x 32 columns
            float s = 0.f;
#pragma unroll
            for (int wv = 0; wv < 8; wv++) s += red[wv][mm][l];
            part[mb + mm][l] = s;
        }
    }
    cluster.sync();
    // rank r finishes the batch rows m with m % nrank == r
    for (int idx = threadIdx.x; idx < MB * 32; idx += 256) {
        const int m = idx >> 5, l = idx & 31;
        const int ii = blockIdx.x * 32 + l;
        if ((m % nrank) == rank && m < M && ii < I) {
            float s = 0.f;
            for (int r = 0; r < nrank; r++) s += cluster.map_shared_rank(&part[0][0], r)[m * 32 + l];
            dx[(long long)m * I + ii] = s * wgain;
        }
    }
    cluster.sync();        // keep this CTA's partial sums alive until every rank has read them
}

// a thread owns one input feature (x column in registers), a CTA 16 output features
template <int MB>
__global__ void __launch_bounds__(256) fc_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                       float* __restrict__ db, int M, int I, int O, float wgain, float bgain) {
    constexpr int OB = 16;
    __shared__ float dys[OB][MB];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int o0 = blockIdx.y * OB;
    for (int idx = threadIdx.x; idx < OB * MB; idx += 256) {
        const int m = idx / OB, oo = idx - m * OB;
        dys[oo][m] = (m < M && o0 + oo < O) ? dy[(long long)m * O + o0 + oo] : 0.f;
    }
    float xr[MB];
#pragma unroll
    for (int m = 0; m < MB; m++) xr[m] = (m < M && i < I) ? x[(long long)m * I + i] : 0.f;
    __syncthreads();
    if (i < I) {
#pragma unroll 4
        for (int oo = 0; oo < OB; oo++) {
            if (o0 + oo < O) {
                float s = 0.f;
#pragma unroll
                for (int m = 0; m < MB; m++) s += dys[oo][m] * xr[m];
                dw[(long long)(o0 + oo) * I + i] = s * wgain;
            }
        }
    }
    if (db && blockIdx.x == 0 && threadIdx.x < OB && o0 + threadIdx.x < O) {
        float s = 0.f;
        for (int m = 0; m < MB; m++) s += dys[threadIdx.x][m];
        db[o0 + threadIdx.x] = s * bgain;
    }
}

inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// launch with a (1, cluster_y, 1) thread-block cluster
template <class... KArgs, class... Args>
cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, int cluster_y, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = cluster_y;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace

#define GT_FC_MB(M_, CALL)                 \
    if (M_ <= 8) { CALL(8) }               \
    else if (M_ <= 16) { CALL(16) }        \
    else if (M_ <= 32) { CALL(32) }        \
    else { CALL(64) }

extern "C" int gt_fc_fwd(const float* x, const float* w, const float* b, float* y, int M, int I, int O, float wgain, float bgain, void* stream) {
    GT_REQUIRE(x && w && y, "gt_fc_fwd: null pointer");
    GT_REQUIRE(M >= 1 && M <= 64 && I >= 1 && O >= 1, "gt_fc_fwd: shape M=%d I=%d O=%d not supported (1 <= M <= 64)", M, I, O);
    GT_REQUIRE(I % 4 == 0 && al16(x) && al16(w), "gt_fc_fwd: in_features must be a multiple of 4 and x, w 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // k-slices per cluster: enough CTAs to cover the SMs, at least one 128-wide chunk per slice
    const int cols = (O + 7) / 8, chunks = (I + FC_KC - 1) / FC_KC;
    int ks = 1;
    while (ks < 8 && cols * ks * 2 <= 2 * gt_num_sms() && ks * 2 <= chunks) ks *= 2;
    const int k_per_rank = ((chunks + ks - 1) / ks) * FC_KC;
    cudaError_t le = cudaSuccess;
#define CALL(MB_) le = launch_cluster(fc_fwd_kernel<MB_>, dim3(cols, ks), ks, st, x, w, b, y, M, I, O, wgain, bgain, k_per_rank);
    GT_FC_MB(M, CALL)
#undef CALL
    if (le != cudaSuccess) {
        gt_set_error("gt_fc_fwd: launch failed: %s", cudaGetErrorString(le));
        return GT_ERR_CUDA;
    }
    return GT_OK;
}

extern "C" int gt_fc_dgrad(const float* dy, const float* w, float* dx, int M, int I, int O, float wgain, void* stream) {
    GT_REQUIRE(dy && w && dx, "gt_fc_dgrad: null pointer");
    GT_REQUIRE(M >= 1 && M <= 64 && I >= 1 && O >= 1, "gt_fc_dgrad: shape M=%d I=%d O=%d not supported (1 <= M <= 64)", M, I, O);
    cudaStream_t st = (cudaStream_t)stream;
    // output-feature slices per cluster: enough CTAs to cover the SMs, at least 64 features per slice
    const int cols = (I + 31) / 32;
    int os = 1;
    while (os < 8 && cols * os * 2 <= 2 * gt_num_sms() && os * 2 * 64 <= O) os *= 2;
    const int o_per_rank = (O + os - 1) / os;
    cudaError_t le = cudaSuccess;
#define CALL(MB_) le = launch_cluster(fc_dgrad_kernel<MB_>, dim3(cols, os), os, st, dy, w, dx, M, I, O, wgain, o_per_rank);
    GT_FC_MB(M, CALL)
#undef CALL
    if (le != cudaSuccess) {
        gt_set_error("gt_fc_dgrad: launch failed: %s", cudaGetErrorString(le));
        return GT_ERR_CUDA;
    }
    return GT_OK;
}

extern "C" int gt_fc_wgrad(const float* dy, const float* x, float* dw, float* db, int M, int I, int O, float wgain, float bgain, void* stream) {
    GT_REQUIRE(dy && x && dw, "gt_fc_wgrad: null pointer");
    GT_REQUIRE(M >= 1 && M <= 64 && I >= 1 && O >= 1, "gt_fc_wgrad: shape M=%d I=%d O=%d not supported (1 <= M <= 64)", M, I, O);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((I + 255) / 256, (O + 15) / 16);
#define CALL(MB_) fc_wgrad_kernel<MB_><<<grid, 256, 0, st>>>(dy, x, dw, db, M, I, O, wgain, bgain);
    GT_FC_MB(M, CALL)
#undef CALL
    GT_CUDA_LAUNCH_CHECK("gt_fc_wgrad");
    return GT_OK;
}
