// conv_wgrad_halo.cu -- weight gradient of the stride-1 3x3 fp16 NHWC convolutions: halo-staged, tap-paired split-K
// implicit GEMM on tcgen05 + TMEM + TMA.  Same contract as conv_wgrad.cu (which keeps the strided / transposed-stride-2 /
// 1x1 / tiny-image cases); this kernel exists because the per-tap-row kernel is bound by L2 -> SM traffic on the layers
// that carry the pixels (ncu, profiles/r01b_ncu_full_summary.txt: 4.0 GB through the crossbar for 0.54 GB of algorithmic
// traffic on the 64-channel 256x256 layer -- each of the three kernel rows re-reads U, each tap re-reads S -- and half
// of every MMA is zero padding when there are only 64 U channels).
//
//     G[u_ch][s_ch][r][s] = sum_{n,i,j} U[n,i,j,u_ch] * S[n, i + r - pad, j + s - pad, s_ch]          (conv_wgrad.cu, stride 1)
//
// What changes:
//   * ONE TMA box per pixel tile brings the S footprint of ALL nine taps ((8 + 2) x (16 + 2) pixels x 64 channels); a
//     tap is a start address inside it (+ (r * pitch + s) * 128 bytes; the 128-byte swizzle is a function of the address
//     bits, so any 128-byte row may start an operand), the 8-pixel tile rows are the 8-row groups of the MN-major
//     operand and the halo row pitch is the descriptor's stride byte offset.
//   * Taps are PAIRED into the M dimension: the MN-major A operand is two 64-channel groups `LBO` bytes apart, and with
//     LBO = (address of tap B) - (address of tap A) the second group is the same staged tile seen through the other tap:
//         A[m][k] = S[pixel k + shift(tap_A)][s_ch m]         m <  64
//                 = S[pixel k + shift(tap_B)][s_ch m - 64]    m >= 64
//     B = U tile (N = 64 U channels, MN-major, dense).  One M = 128, N = 64 MMA therefore produces two taps; nine taps
//     are five pairs (the ninth pairs with itself), 5 x 64 = 320 accumulator columns -- all nine taps of a 64 x 64
//     channel tile stay in TMEM for the whole pixel range, and no MMA row is padding.
//   * A CTA owns one (64 S-channel, 64 U-channel) tile and a slice of the pixel tiles (split-K); fp32 partials go to the
//     workspace [split][tap][u][s] and wgrad_reduce sums them in split order (deterministic, conv_wgrad.cu).
// Pipeline: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue; STAGES x {U 16 KB, S 23 KB}.
#include "gt_common.cuh"
#include "gt_sm100.cuh"

using namespace sm100;

extern thread_local int t_wgrad_px_limit;      // conv_wgrad.cu: > 0 = at most this many pixels per split-K slice (fp16x3 route)

namespace {

constexpr int NTHREADS = 192;
constexpr int TW = 8, TH = 16;                 // pixel tile (8-pixel rows = one MN-major 8-row group)
constexpr int HW_ = TW + 2, HH_ = TH + 2;      // staged S footprint for 3x3
constexpr int STAGES = 4;
constexpr uint32_t U_BYTES = TW * TH * 128;                                   // 16 KB
constexpr uint32_t S_BYTES = ((HW_ * HH_ * 128 + 1023) / 1024) * 1024;        // 23 KB (23040 used)
constexpr uint32_t STAGE_BYTES = U_BYTES + S_BYTES;
constexpr uint32_t SMEM_TOTAL = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 8 + 1024;
constexpr int NPAIRS = 5;
constexpr uint32_t TMEM_COLS = 512;            // 5 x 64 accumulator columns, rounded up to a power of two

struct WHParams {
    int N, UH, UW, UC, SC;
    int pad;
    int tiles_w, tiles_h, num_tiles, splits, s_tiles;
    float* ws;                 // [splits][9][UC][SC]
};

__global__ void __launch_bounds__(NTHREADS, 1) conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmS,
                                                                      const WHParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = full + 2 * STAGES;
    uint32_t* tmem_slot = (uint32_t*)(full + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int ut = blockIdx.y / p.s_tiles, st = blockIdx.y % p.s_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmU);
        tma_prefetch_desc(&tmS);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = split; t < p.num_tiles; t += p.splits) {
                const int twi = t % p.tiles_w;
                const int rest = t / p.tiles_w;
                const int thi = rest % p.tiles_h, n = rest / p.tiles_h;
                const int j0 = twi * TW, i0 = thi * TH;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], U_BYTES + HW_ * HH_ * 128);
                uint8_t* sU = smem + stage * STAGE_BYTES;
                tma_load_4d(sU, &tmU, &full[stage], ut * 64, j0, i0, n);                                   // rows/cols past the image read as zero
                tma_load_4d(sU + U_BYTES, &tmS, &full[stage], st * 64, j0 - p.pad, i0 - p.pad, n);         // = the convolution's zero padding
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(128, 64, 0, 1, 1);   // both operands MN-major
            constexpr uint32_t pitch = HW_ * 128;                      // bytes between tile rows of the staged S footprint
            int stage = 0;
            uint32_t phase = 0;
            bool first = true;
            for (int t = split; t < p.num_tiles; t += p.splits) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t u0 = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t s0 = u0 + U_BYTES;
#pragma unroll
                for (int pr = 0; pr < NPAIRS; pr++) {
                    const int ta = 2 * pr, tb = (2 * pr + 1 < 9) ? 2 * pr + 1 : 2 * pr;
                    const uint32_t offa = (uint32_t)((ta / 3) * HW_ + (ta % 3)) * 128u, offb = (uint32_t)((tb / 3) * HW_ + (tb % 3)) * 128u;
#pragma unroll
                    for (int k = 0; k < TH / 2; k++)      // K = 16 pixels = two 8-pixel tile rows per MMA
                        umma_f16(tmem_base + pr * 64, umma_smem_desc(s0 + offa + (uint32_t)(2 * k) * pitch, offb - offa, pitch),
                                 umma_smem_desc(u0 + (uint32_t)(2 * k) * 1024u, 0, 1024), idesc, (uint32_t)(!first || k != 0));
                }
                first = false;
                umma_commit(&empty[stage]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(tfull);
        }
        __syncwarp();
    } else {
        // TMEM lane m = (pair member, S channel): warps with (warp & 3) in {0, 1} hold tap A of every pair, {2, 3} tap B
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int member = m >> 6;
        const int s = st * 64 + (m & 63);
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int pr = 0; pr < NPAIRS; pr++) {
            const int tap = 2 * pr + member;
            const bool valid = tap < 9;                       // the ninth tap is paired with itself: drop the copy
            float* wp = p.ws + (((long long)split * 9 + (valid ? tap : 0)) * p.UC + ut * 64) * p.SC + s;
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pr * 64 + c * 32), r);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int v = 0; v < 32; v++) wp[(long long)(c * 32 + v) * p.SC] = __uint_as_float(r[v]);   // a warp writes 32 consecutive s
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

struct WHPlan {
    int tiles_w, tiles_h, num_tiles, splits, u_tiles, s_tiles;
};

WHPlan make_plan(int N, int UH, int UW, int UC, int SC) {
    WHPlan pl;
    pl.tiles_w = (UW + TW - 1) / TW;
    pl.tiles_h = (UH + TH - 1) / TH;
    pl.num_tiles = pl.tiles_w * pl.tiles_h * N;
    pl.u_tiles = UC / 64;
    pl.s_tiles = SC / 64;
    const int col_tiles = pl.u_tiles * pl.s_tiles;
    int splits = gt_num_sms() / col_tiles;          // one CTA per SM (the accumulators take the whole TMEM)
    if (splits > pl.num_tiles / 4) splits = pl.num_tiles / 4;   // at least four pixel tiles per CTA to amortise the 9-tap epilogue
    if (t_wgrad_px_limit > 0) {                     // fp16x3 route: bound the accumulator updates per split-K slice (csrc/conv_f16x3.cu)
        const long long px = (long long)N * UH * UW;
        const int need = (int)((px + t_wgrad_px_limit - 1) / t_wgrad_px_limit);
        if (splits < need) splits = need;
        if (splits > pl.num_tiles) splits = pl.num_tiles;
    }
    if (splits < 1) splits = 1;
    pl.splits = splits;
    return pl;
}

}  // namespace

bool gt_wgrad_halo_applicable(int N, int UH, int UW, int UC, int SC, int SH, int SW, int KH, int KW, int stride, int pad) {
    if (KH != 3 || KW != 3 || stride != 1 || pad < 0 || pad > 2) return false;
    if (UC % 64 || SC % 64 || UH < TH || UW < TW || N < 1) return false;
    if (SH != UH + 2 - 2 * pad || SW != UW + 2 - 2 * pad) return false;     // U = conv output of S (or S = transposed-conv output of U)
    // worth it when the pixels dominate: enough tiles per CTA to amortise the 9-tap epilogue
    WHPlan pl = make_plan(N, UH, UW, UC, SC);
    return pl.num_tiles >= 32;
}

long long gt_wgrad_halo_workspace(int N, int UH, int UW, int UC, int SC) {
    WHPlan pl = make_plan(N, UH, UW, UC, SC);
    return (long long)pl.splits * 9 * UC * SC;
}

// returns the number of splits written to the workspace (> 0) or a negative error indicator after gt_set_error
int gt_launch_wgrad_halo(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s, long long ss_n, long long ss_h,
                         long long ss_w, int SH, int SW, int SC, int N, int pad, float* workspace, long long workspace_floats, cudaStream_t stream) {
    WHPlan pl = make_plan(N, UH, UW, UC, SC);
    if (workspace_floats < (long long)pl.splits * 9 * UC * SC) {
        gt_set_error("gt_conv2d_wgrad_f16 (halo): workspace too small");
        return -1;
    }
    gt_encode_tiled_fn encode = gt_get_encode_tiled();
    if (!encode) {
        gt_set_error("gt_conv2d_wgrad_f16: cuTensorMapEncodeTiled is not available from this driver");
        return -1;
    }
    CUtensorMap tmU, tmS;
    {
        cuuint64_t dims[4] = {(cuuint64_t)UC, (cuuint64_t)UW, (cuuint64_t)UH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)us_w * 2, (cuuint64_t)us_h * 2, (cuuint64_t)us_n * 2};
        cuuint32_t box[4] = {64, TW, TH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmU, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(u), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            gt_set_error("gt_conv2d_wgrad_f16 (halo): U tensor map rejected (CUresult %d)", (int)r);
            return -1;
        }
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)SC, (cuuint64_t)SW, (cuuint64_t)SH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ss_w * 2, (cuuint64_t)ss_h * 2, (cuuint64_t)ss_n * 2};
        cuuint32_t box[4] = {64, HW_, HH_, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            gt_set_error("gt_conv2d_wgrad_f16 (halo): S tensor map rejected (CUresult %d)", (int)r);
            return -1;
        }
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL);
        if (e != cudaSuccess) {
            gt_set_error("gt_conv2d_wgrad_f16 (halo): cannot reserve %u bytes of shared memory: %s", SMEM_TOTAL, cudaGetErrorString(e));
            return -1;
        }
        configured = true;
    }
    WHParams p;
    memset(&p, 0, sizeof(p));
    p.N = N;
    p.UH = UH;
    p.UW = UW;
    p.UC = UC;
    p.SC = SC;
    p.pad = pad;
    p.tiles_w = pl.tiles_w;
    p.tiles_h = pl.tiles_h;
    p.num_tiles = pl.num_tiles;
    p.splits = pl.splits;
    p.s_tiles = pl.s_tiles;
    p.ws = workspace;
    dim3 grid((unsigned)pl.splits, (unsigned)(pl.u_tiles * pl.s_tiles), 1);
    conv_wgrad_halo_kernel<<<grid, NTHREADS, SMEM_TOTAL, stream>>>(tmU, tmS, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        gt_set_error("gt_conv2d_wgrad_f16 (halo): CUDA launch failed: %s", cudaGetErrorString(e));
        return -1;
    }
    return pl.splits;
}
