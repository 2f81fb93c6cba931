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
//
// STRIDE = 2 (weight gradient of the discriminator's stride-2 down-convolutions and of the generator's transposed stride-2
// up-convolutions, pad 0):  S is read at (2 i + r, 2 j + s).  The footprint is staged as its four PARITY PLANES
// P[pr][ps][y][x] = S[2 (i0 + y) + pr][2 (j0 + x) + ps] -- four TMA boxes of (8 + 1) x (8 + 1) pixels with element strides 2 -- so that
// tap (r, s) is again a dense window: plane (r & 1, s & 1) shifted by (r >> 1, s >> 1).  The taps are paired in the order of their
// byte offsets inside the stage (the descriptor's LBO is unsigned); everything else is the stride-1 kernel with an 8 x 8 pixel tile.
// The per-tap-row kernel these cases used before fetched every S tile once per tap and was bound by L2 -> SM traffic
// (137-189 us against 61-139 us for the library at the training shapes, profiles/r02n_evidence_call_stdout.txt).
#include "gt_common.cuh"
#include "gt_sm100.cuh"

using namespace sm100;

extern thread_local int t_wgrad_px_limit;      // conv_wgrad.cu: > 0 = at most this many pixels per split-K slice (fp16x3 route)

namespace {

constexpr int NTHREADS = 192;
constexpr int TW = 8;                          // pixel tile width (8-pixel rows = one MN-major 8-row group)
constexpr int STAGES = 4;
constexpr int NPAIRS = 5;
constexpr uint32_t TMEM_COLS = 512;            // 5 x 64 accumulator columns, rounded up to a power of two

// geometry of one pipeline stage: {U tile | staged S footprint}
template <int STRIDE>
struct Geo;
template <>
struct Geo<1> {
    static constexpr int TH = 16;                                                  // pixel tile height
    static constexpr int PW = TW + 2, PH = TH + 2;                                 // staged S footprint for 3x3
    static constexpr int NPLANES = 1;
    static constexpr uint32_t PLANE_BYTES = ((PW * PH * 128 + 1023) / 1024) * 1024;   // 23 KB (23040 used)
    static constexpr uint32_t PLANE_TX = PW * PH * 128;
    // tap t = r * 3 + s in pairing order, and its byte offset inside the staged footprint
    __host__ __device__ static constexpr int tap_of(int i) { return i; }
    __host__ __device__ static constexpr uint32_t off_of(int tap) { return (uint32_t)((tap / 3) * PW + (tap % 3)) * 128u; }
};
template <>
struct Geo<2> {
    static constexpr int TH = 8;
    static constexpr int PW = TW + 1, PH = TH + 1;                                 // one parity plane: 9 x 9 pixels
    static constexpr int NPLANES = 4;
    static constexpr uint32_t PLANE_BYTES = ((PW * PH * 128 + 1023) / 1024) * 1024;   // 11 KB (10368 used)
    static constexpr uint32_t PLANE_TX = PW * PH * 128;
    __host__ __device__ static constexpr uint32_t off_of(int tap) {
        return (uint32_t)((((tap / 3) & 1) * 2 + ((tap % 3) & 1))) * PLANE_BYTES + (uint32_t)(((tap / 3) >> 1) * PW + ((tap % 3) >> 1)) * 128u;
    }
    // taps sorted by off_of(): (0,0) (0,2) (2,0) (2,2) | (0,1) (2,1) | (1,0) (1,2) | (1,1)
    __host__ __device__ static constexpr int tap_of(int i) {
        return i == 0 ? 0 : i == 1 ? 2 : i == 2 ? 6 : i == 3 ? 8 : i == 4 ? 1 : i == 5 ? 7 : i == 6 ? 3 : i == 7 ? 5 : 4;
    }
};
template <int STRIDE>
struct StageLayout {
    typedef Geo<STRIDE> G;
    static constexpr uint32_t U_BYTES = TW * G::TH * 128;
    static constexpr uint32_t S_BYTES = G::NPLANES * G::PLANE_BYTES;
    static constexpr uint32_t STAGE_BYTES = U_BYTES + S_BYTES;
    static constexpr uint32_t TX_BYTES = U_BYTES + G::NPLANES * G::PLANE_TX;
    static constexpr uint32_t SMEM_TOTAL = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 8 + 1024;
};
static_assert(Geo<2>::off_of(0) < Geo<2>::off_of(2) && Geo<2>::off_of(2) < Geo<2>::off_of(6) && Geo<2>::off_of(6) < Geo<2>::off_of(8) &&
                  Geo<2>::off_of(8) < Geo<2>::off_of(1) && Geo<2>::off_of(1) < Geo<2>::off_of(7) && Geo<2>::off_of(7) < Geo<2>::off_of(3) &&
                  Geo<2>::off_of(3) < Geo<2>::off_of(5) && Geo<2>::off_of(5) < Geo<2>::off_of(4),
              "stride-2 taps must be paired in increasing offset order");
static_assert(StageLayout<2>::SMEM_TOTAL <= 232448 && StageLayout<1>::SMEM_TOTAL <= 232448, "stage ring exceeds the shared memory of one SM");

struct WHParams {
    int N, UH, UW, UC, SC;
    int pad;
    int tiles_w, tiles_h, num_tiles, splits, s_tiles;
    float* ws;                 // [splits][9][UC][SC]
};

template <int STRIDE>
__global__ void __launch_bounds__(NTHREADS, 1) conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmS,
                                                                      const WHParams p) {
    typedef Geo<STRIDE> G;
    typedef StageLayout<STRIDE> L;
    constexpr int TH = G::TH;
    constexpr uint32_t U_BYTES = L::U_BYTES, STAGE_BYTES = L::STAGE_BYTES;
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
                mbar_arrive_expect_tx(&full[stage], L::TX_BYTES);
                uint8_t* sU = smem + stage * STAGE_BYTES;
                tma_load_4d(sU, &tmU, &full[stage], ut * 64, j0, i0, n);                                   // rows/cols past the image read as zero
                if (STRIDE == 1) {
                    tma_load_4d(sU + U_BYTES, &tmS, &full[stage], st * 64, j0 - p.pad, i0 - p.pad, n);     // = the convolution's zero padding
                } else {
#pragma unroll
                    for (int pl = 0; pl < 4; pl++)                                                         // parity plane (pr, ps) = (pl >> 1, pl & 1)
                        tma_load_4d(sU + U_BYTES + pl * G::PLANE_BYTES, &tmS, &full[stage], st * 64, 2 * j0 + (pl & 1), 2 * i0 + (pl >> 1), n);
                }
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
            constexpr uint32_t pitch = G::PW * 128;                    // bytes between tile rows of the staged S footprint
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
                    const uint32_t offa = G::off_of(G::tap_of(2 * pr)), offb = G::off_of(G::tap_of((2 * pr + 1 < 9) ? 2 * pr + 1 : 2 * pr));
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
            const bool valid = 2 * pr + member < 9;           // the ninth tap is paired with itself: drop the copy
            const int tap = G::tap_of(valid ? 2 * pr + member : 8);
            float* wp = p.ws + (((long long)split * 9 + tap) * p.UC + ut * 64) * p.SC + s;
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

WHPlan make_plan(int N, int UH, int UW, int UC, int SC, int stride) {
    const int TH = stride == 1 ? Geo<1>::TH : Geo<2>::TH;
    WHPlan pl;
    pl.tiles_w = (UW + TW - 1) / TW;
    pl.tiles_h = (UH + TH - 1) / TH;
    pl.num_tiles = pl.tiles_w * pl.tiles_h * N;
    pl.u_tiles = UC / 64;
    pl.s_tiles = SC / 64;
    const int col_tiles = pl.u_tiles * pl.s_tiles;
    const int min_tiles = stride == 1 ? 4 : 8;      // pixel tiles per CTA that amortise the 9-tap epilogue (128 / 64 pixels per tile)
    int splits = gt_num_sms() / col_tiles;          // one CTA per SM (the accumulators take the whole TMEM)
    if (splits > pl.num_tiles / min_tiles) splits = pl.num_tiles / min_tiles;
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

template <int STRIDE>
int launch(const CUtensorMap& tmU, const CUtensorMap& tmS, const WHParams& p, const WHPlan& pl, cudaStream_t stream) {
    typedef StageLayout<STRIDE> L;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo_kernel<STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM_TOTAL);
        if (e != cudaSuccess) {
            gt_set_error("gt_conv2d_wgrad_f16 (halo): cannot reserve %u bytes of shared memory: %s", L::SMEM_TOTAL, cudaGetErrorString(e));
            return -1;
        }
        configured = true;
    }
    dim3 grid((unsigned)pl.splits, (unsigned)(pl.u_tiles * pl.s_tiles), 1);
    conv_wgrad_halo_kernel<STRIDE><<<grid, NTHREADS, L::SMEM_TOTAL, stream>>>(tmU, tmS, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        gt_set_error("gt_conv2d_wgrad_f16 (halo): CUDA launch failed: %s", cudaGetErrorString(e));
        return -1;
    }
    return pl.splits;
}

}  // namespace

static int g_wgrad_halo_s2 = 1;        // 0: strided cases stay on the per-tap-row kernel (A/B switch, gt_conv_wgrad_config(2))
void gt_wgrad_halo_enable_s2(int on) { g_wgrad_halo_s2 = on; }

bool gt_wgrad_halo_applicable(int N, int UH, int UW, int UC, int SC, int SH, int SW, int KH, int KW, int stride, int pad) {
    if (KH != 3 || KW != 3 || pad < 0 || N < 1 || UC % 64 || SC % 64) return false;
    if (stride == 1) {
        if (pad > 2 || UH < Geo<1>::TH || UW < TW) return false;
        if (SH != UH + 2 - 2 * pad || SW != UW + 2 - 2 * pad) return false;     // U = conv output of S (or S = transposed-conv output of U)
    } else if (stride == 2) {
        // S is read at (2 i + r, 2 j + s), r, s in 0..2: the pad-0 down-convolution (S = x, 2 UH + 1 or 2 UH + 2 rows) and the pad-0
        // transposed up-convolution (S = dy, 2 UH + 1 rows); rows / columns past S read as zero either way
        if (!g_wgrad_halo_s2 || pad != 0 || UH < Geo<2>::TH || UW < TW) return false;
        if (SH < 2 * UH + 1 || SH > 2 * UH + 2 || SW < 2 * UW + 1 || SW > 2 * UW + 2) return false;
    } else {
        return false;
    }
    // worth it when the pixels dominate: enough tiles per CTA to amortise the 9-tap epilogue
    WHPlan pl = make_plan(N, UH, UW, UC, SC, stride);
    return pl.num_tiles >= (stride == 1 ? 32 : 64);
}

long long gt_wgrad_halo_workspace(int N, int UH, int UW, int UC, int SC, int stride) {
    WHPlan pl = make_plan(N, UH, UW, UC, SC, stride);
    return (long long)pl.splits * 9 * UC * SC;
}

// returns the number of splits written to the workspace (> 0) or a negative error indicator after gt_set_error
int gt_launch_wgrad_halo(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s, long long ss_n, long long ss_h,
                         long long ss_w, int SH, int SW, int SC, int N, int stride, int pad, float* workspace, long long workspace_floats,
                         cudaStream_t stream) {
    WHPlan pl = make_plan(N, UH, UW, UC, SC, stride);
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
        cuuint32_t box[4] = {64, TW, (cuuint32_t)(stride == 1 ? Geo<1>::TH : Geo<2>::TH), 1};
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
        // stride 1: the whole 10 x 18 pixel footprint; stride 2: one 9 x 9 pixel parity plane per load (every other pixel / row of an
        // 18 x 18 window: with element strides the box extent counts tensor elements, ceil(18 / 2) = 9 of them are written)
        cuuint32_t box[4] = {64, (cuuint32_t)(stride == 1 ? Geo<1>::PW : 2 * Geo<2>::PW), (cuuint32_t)(stride == 1 ? Geo<1>::PH : 2 * Geo<2>::PH), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = encode(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            gt_set_error("gt_conv2d_wgrad_f16 (halo): S tensor map rejected (CUresult %d)", (int)r);
            return -1;
        }
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
    return stride == 1 ? launch<1>(tmU, tmS, p, pl, stream) : launch<2>(tmU, tmS, p, pl, stream);
}
