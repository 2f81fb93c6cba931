// conv_wgrad_halo_wide.cu -- weight gradient of the stride-1 3x3 fp16 NHWC convolutions with >= 128 U channels: the halo-staged,
// tap-paired kernel of conv_wgrad_halo.cu with N = 128 U channels per CTA and the nine taps split over TWO CTA TYPES.
//
// Why.  Measured MMA pacing (tools/umma_probe*.cu, profiles/r02_umma_probe.txt): an SS-mode tcgen05.mma M128 x N x K16 takes
// max(N / 2, 32 + N / 4) clocks, so the N = 64 pair-MMAs of conv_wgrad_halo.cu run at 2/3 of the array rate and, with five pairs for
// nine taps, the kernel stops at 0.6 of it (ncu: 59 % tensor-pipe activity).  N = 128 reaches the array rate, but nine resident taps
// x 64 S channels x 128 U channels of fp32 accumulators do not fit the 512 TMEM columns (four pairs x 128 columns hold eight taps).
// So the kernel rows are split:
//     type A CTAs keep rows {0, 1} = taps 0..5 = three pairs  (384 columns, 3 MMAs of N = 128 per K step: all rows useful)
//     type B CTAs keep row  {2}    = taps 6..8 = two pairs    (256 columns, 2 MMAs per K step, the ninth tap pairs with itself)
// and the pixel range of a (64 S-channel, 128 U-channel) tile is cut into nA slices for type A and nB slices for type B with
// nA : nB ~ 3 : 2, so that both types finish together: 320 clocks of MMA per K step and 128 U channels instead of 480 -- 0.9 of the
// array rate.  Both types stage the same {U 2 x 16 KB, S 23 KB} per pixel tile (the S halo box serves every tap, the U tile both
// 64-channel groups).  Partial sums go to two workspace regions ([nA][6 taps][UC][SC], [nB][3 taps][UC][SC]) that
// wgrad_reduce_ab (conv_wgrad.cu) sums in slice order -> deterministic, as before.
//
// STRIDE = 2: the parity-plane staging of conv_wgrad_halo.cu (four 9 x 9 pixel planes per 8 x 8 pixel tile, taps in offset order
// 0 2 6 8 | 1 7 | 3 5 | 4): type A keeps the first six taps of that order, type B the last three.
#include "gt_common.cuh"
#include "gt_sm100.cuh"

using namespace sm100;

extern thread_local int t_wgrad_px_limit;      // conv_wgrad.cu: > 0 = at most this many pixels per split-K slice (fp16x3 route)

namespace {

constexpr int NTHREADS = 192;
constexpr int TW = 8;                          // pixel tile width

template <int STRIDE>
struct Geo;
template <>
struct Geo<1> {
    static constexpr int TH = 16, PW = TW + 2, PH = TH + 2, NPLANES = 1, STAGES = 4;
    static constexpr uint32_t PLANE_BYTES = ((PW * PH * 128 + 1023) / 1024) * 1024;      // 23 KB (23040 used)
    __host__ __device__ static constexpr int tap_of(int i) { return i; }
    __host__ __device__ static constexpr uint32_t off_of(int tap) { return (uint32_t)((tap / 3) * PW + (tap % 3)) * 128u; }
};
template <>
struct Geo<2> {
    static constexpr int TH = 8, PW = TW + 1, PH = TH + 1, NPLANES = 4, STAGES = 3;
    static constexpr uint32_t PLANE_BYTES = ((PW * PH * 128 + 1023) / 1024) * 1024;      // 11 KB (10368 used)
    __host__ __device__ static constexpr uint32_t off_of(int tap) {
        return (uint32_t)((((tap / 3) & 1) * 2 + ((tap % 3) & 1))) * PLANE_BYTES + (uint32_t)(((tap / 3) >> 1) * PW + ((tap % 3) >> 1)) * 128u;
    }
    __host__ __device__ static constexpr int tap_of(int i) {          // taps sorted by off_of()
        return i == 0 ? 0 : i == 1 ? 2 : i == 2 ? 6 : i == 3 ? 8 : i == 4 ? 1 : i == 5 ? 7 : i == 6 ? 3 : i == 7 ? 5 : 4;
    }
};
template <int STRIDE>
struct Lay {
    typedef Geo<STRIDE> G;
    static constexpr int STAGES = G::STAGES;
    static constexpr uint32_t UBOX = TW * G::TH * 128;                                // one 64-channel U box
    static constexpr uint32_t U_BYTES = 2 * UBOX;                                     // 128 U channels
    static constexpr uint32_t S_BYTES = G::NPLANES * G::PLANE_BYTES;
    static constexpr uint32_t STAGE_BYTES = U_BYTES + S_BYTES;
    static constexpr uint32_t TX_BYTES = U_BYTES + G::NPLANES * G::PW * G::PH * 128;
    static constexpr uint32_t SMEM_TOTAL = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 8 + 1024;
};
static_assert(Lay<1>::SMEM_TOTAL <= 232448 && Lay<2>::SMEM_TOTAL <= 232448, "stage ring exceeds the shared memory of one SM");

struct WWParams {
    int N, UH, UW, UC, SC;
    int pad;
    int tiles_w, tiles_h, num_tiles, s_tiles;
    int nA, nB;                // pixel slices per channel tile for the two CTA types
    float* wsA;                // [nA][6][UC][SC]
    float* wsB;                // [nB][3][UC][SC]
};

// NP pairs starting at position T0 of the tap order (the last pair of type B pairs the ninth tap with itself)
template <int STRIDE, int NP, int T0>
__device__ __forceinline__ void issue_tiles(uint8_t* smem, uint64_t* full, uint64_t* empty, uint64_t* tfull, uint32_t tmem_base, int slice, int nslices,
                                            int num_tiles) {
    typedef Geo<STRIDE> G;
    typedef Lay<STRIDE> L;
    constexpr int STAGES = L::STAGES, TH = G::TH;
    constexpr uint32_t STAGE_BYTES = L::STAGE_BYTES, U_BYTES = L::U_BYTES, UBOX = L::UBOX;
    constexpr uint32_t idesc = umma_idesc(128, 128, 0, 1, 1);   // both operands MN-major
    constexpr uint32_t pitch = G::PW * 128;
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int t = slice; t < num_tiles; t += nslices) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t u0 = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t s0 = u0 + U_BYTES;
#pragma unroll
        for (int pr = 0; pr < NP; pr++) {
            const int pa = T0 + 2 * pr, pb = (pa + 1 < 9) ? pa + 1 : pa;
            const uint32_t offa = G::off_of(G::tap_of(pa)), offb = G::off_of(G::tap_of(pb));
#pragma unroll
            for (int k = 0; k < TH / 2; k++)      // K = 16 pixels = two 8-pixel tile rows per MMA
                umma_f16(tmem_base + pr * 128, umma_smem_desc(s0 + offa + (uint32_t)(2 * k) * pitch, offb - offa, pitch),
                         umma_smem_desc(u0 + (uint32_t)(2 * k) * 1024u, UBOX, 1024), idesc, (uint32_t)(!first || k != 0));
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

// TMEM lane m = (pair member, S channel); columns = 128 U channels of pair pr.  ws: [slice][NT taps][UC][SC]
template <int NP, int T0, int NT>
__device__ __forceinline__ void drain(float* ws, int slice, int UC, int SC, int ut, int st, uint32_t tmem_base, int warp, int lane) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int member = m >> 6;
    const int s = st * 64 + (m & 63);
#pragma unroll 1
    for (int pr = 0; pr < NP; pr++) {
        const int tl = 2 * pr + member;                  // tap index local to this CTA type
        const bool valid = T0 + tl < 9;                  // the ninth tap is paired with itself: drop the copy
        float* wp = ws + (((long long)slice * NT + (valid ? tl : 0)) * UC + ut * 128) * SC + s;
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pr * 128 + c * 32), r);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int v = 0; v < 32; v++) wp[(long long)(c * 32 + v) * SC] = __uint_as_float(r[v]);   // a warp writes 32 consecutive s
            }
        }
    }
}

template <int STRIDE>
__global__ void __launch_bounds__(NTHREADS, 1) conv_wgrad_halo_wide_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmS,
                                                                           const WWParams p) {
    typedef Geo<STRIDE> G;
    typedef Lay<STRIDE> L;
    constexpr int STAGES = L::STAGES, TH = G::TH;
    constexpr uint32_t STAGE_BYTES = L::STAGE_BYTES, U_BYTES = L::U_BYTES, UBOX = L::UBOX, TX_BYTES = L::TX_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = full + 2 * STAGES;
    uint32_t* tmem_slot = (uint32_t*)(full + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool typeA = (int)blockIdx.x < p.nA;
    const int slice = typeA ? (int)blockIdx.x : (int)blockIdx.x - p.nA;
    const int nslices = typeA ? p.nA : p.nB;
    const int ut = blockIdx.y / p.s_tiles, st = blockIdx.y % p.s_tiles;
    const uint32_t tmem_cols = typeA ? 512u : 256u;

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
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = slice; t < p.num_tiles; t += nslices) {
                const int twi = t % p.tiles_w;
                const int rest = t / p.tiles_w;
                const int thi = rest % p.tiles_h, n = rest / p.tiles_h;
                const int j0 = twi * TW, i0 = thi * TH;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], TX_BYTES);
                uint8_t* sU = smem + stage * STAGE_BYTES;
                tma_load_4d(sU, &tmU, &full[stage], ut * 128, j0, i0, n);                                  // rows/cols past the image read as zero
                tma_load_4d(sU + UBOX, &tmU, &full[stage], ut * 128 + 64, j0, i0, n);
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
            if (typeA) issue_tiles<STRIDE, 3, 0>(smem, full, empty, tfull, tmem_base, slice, nslices, p.num_tiles);
            else issue_tiles<STRIDE, 2, 6>(smem, full, empty, tfull, tmem_base, slice, nslices, p.num_tiles);
        }
        __syncwarp();
    } else {
        mbar_wait(tfull, 0);
        tc_fence_after();
        if (typeA) drain<3, 0, 6>(p.wsA, slice, p.UC, p.SC, ut, st, tmem_base, warp, lane);
        else drain<2, 6, 3>(p.wsB, slice, p.UC, p.SC, ut, st, tmem_base, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

struct WWPlan {
    int tiles_w, tiles_h, num_tiles, u_tiles, s_tiles, nA, nB;
};

WWPlan make_plan(int N, int UH, int UW, int UC, int SC, int stride) {
    const int TH = stride == 1 ? Geo<1>::TH : Geo<2>::TH;
    WWPlan pl;
    pl.tiles_w = (UW + TW - 1) / TW;
    pl.tiles_h = (UH + TH - 1) / TH;
    pl.num_tiles = pl.tiles_w * pl.tiles_h * N;
    pl.u_tiles = UC / 128;
    pl.s_tiles = SC / 64;
    const int col_tiles = pl.u_tiles * pl.s_tiles;
    int per = gt_num_sms() / col_tiles;             // CTAs per channel tile: one CTA per SM (shared memory), one wave
    if (per < 2) per = 2;
    int need = 1;
    if (t_wgrad_px_limit > 0) {                     // fp16x3 route: bound the accumulator updates per split-K slice (csrc/conv_f16x3.cu)
        const long long px = (long long)N * UH * UW;
        need = (int)((px + t_wgrad_px_limit - 1) / t_wgrad_px_limit);
    }
    // type A costs 3 MMAs per K step, type B 2: pick (nA, nB), nA + nB <= per, minimising max(ceil(T / nA) * 3, ceil(T / nB) * 2)
    const int T = pl.num_tiles;
    long long best = -1;
    pl.nA = pl.nB = 1;
    for (int a = 1; a < per; a++) {
        const int b = per - a;
        const long long cost = (long long)((T + a - 1) / a) * 3 > (long long)((T + b - 1) / b) * 2 ? (long long)((T + a - 1) / a) * 3 : (long long)((T + b - 1) / b) * 2;
        if (best < 0 || cost < best) {
            best = cost;
            pl.nA = a;
            pl.nB = b;
        }
    }
    if (pl.nA < need) pl.nA = need;
    if (pl.nB < need) pl.nB = need;
    if (pl.nA > T) pl.nA = T;
    if (pl.nB > T) pl.nB = T;
    return pl;
}

}  // namespace

static int g_wgrad_wide = 1;           // 0: every case stays on the N = 64 kernels (A/B switch, gt_conv_wgrad_config(3))
void gt_wgrad_halo_wide_enable(int on) { g_wgrad_wide = on; }

bool gt_wgrad_halo_wide_applicable(int N, int UH, int UW, int UC, int SC, int SH, int SW, int KH, int KW, int stride, int pad) {
    if (!g_wgrad_wide || KH != 3 || KW != 3 || pad < 0 || N < 1 || UC % 128 || SC % 64) return false;
    if (stride == 1) {
        if (pad > 2 || UH < Geo<1>::TH || UW < TW) return false;
        if (SH != UH + 2 - 2 * pad || SW != UW + 2 - 2 * pad) return false;
    } else if (stride == 2) {
        if (pad != 0 || UH < Geo<2>::TH || UW < TW) return false;
        if (SH < 2 * UH + 1 || SH > 2 * UH + 2 || SW < 2 * UW + 1 || SW > 2 * UW + 2) return false;
        // Both CTA types stage every pixel tile, so a channel tile reads its operands twice.  That is free while the operands stay in the
        // 126 MB L2 (512 -> 256 @32^2: 110 -> 89 us) and costs HBM bandwidth when they do not (128 -> 64 @128^2, 404 MB: 102 -> 119 us
        // measured): the strided layers with few channels and many pixels stay on the N = 64 kernel.
        const long long op_bytes = 2ll * N * ((long long)UH * UW * UC + (long long)SH * SW * SC);
        if (op_bytes > (112ll << 20)) return false;
    } else {
        return false;
    }
    WWPlan pl = make_plan(N, UH, UW, UC, SC, stride);
    // enough pixel tiles per CTA to amortise the epilogue (6 x 64 x 128 fp32 per type-A CTA)
    return pl.num_tiles >= (stride == 1 ? 8 : 16) * (pl.nA > pl.nB ? pl.nA : pl.nB);
}

long long gt_wgrad_halo_wide_workspace(int N, int UH, int UW, int UC, int SC, int stride) {
    WWPlan pl = make_plan(N, UH, UW, UC, SC, stride);
    return ((long long)pl.nA * 6 + (long long)pl.nB * 3) * UC * SC;
}

template <int STRIDE>
static int launch_wide(const CUtensorMap& tmU, const CUtensorMap& tmS, const WWParams& p, const WWPlan& pl, cudaStream_t stream) {
    typedef Lay<STRIDE> L;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo_wide_kernel<STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM_TOTAL);
        if (e != cudaSuccess) {
            gt_set_error("gt_conv2d_wgrad_f16 (wide halo): cannot reserve %u bytes of shared memory: %s", L::SMEM_TOTAL, cudaGetErrorString(e));
            return -1;
        }
        configured = true;
    }
    dim3 grid((unsigned)(pl.nA + pl.nB), (unsigned)(pl.u_tiles * pl.s_tiles), 1);
    conv_wgrad_halo_wide_kernel<STRIDE><<<grid, NTHREADS, L::SMEM_TOTAL, stream>>>(tmU, tmS, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        gt_set_error("gt_conv2d_wgrad_f16 (wide halo): CUDA launch failed: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

// fills nA / nB (the slice counts of the two workspace regions, region B starting nA * 6 * UC * SC floats into the workspace) and pos_of[9]
// (tap r * 3 + s -> its position in the regions: positions 0..5 live in region A, 6..8 in region B); returns 0 or a negative error
// indicator after gt_set_error
int gt_launch_wgrad_halo_wide(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s, long long ss_n,
                              long long ss_h, long long ss_w, int SH, int SW, int SC, int N, int stride, int pad, float* workspace,
                              long long workspace_floats, int* nA, int* nB, int* pos_of, cudaStream_t stream) {
    WWPlan pl = make_plan(N, UH, UW, UC, SC, stride);
    if (workspace_floats < ((long long)pl.nA * 6 + (long long)pl.nB * 3) * UC * SC) {
        gt_set_error("gt_conv2d_wgrad_f16 (wide halo): workspace too small");
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
            gt_set_error("gt_conv2d_wgrad_f16 (wide halo): U tensor map rejected (CUresult %d)", (int)r);
            return -1;
        }
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)SC, (cuuint64_t)SW, (cuuint64_t)SH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ss_w * 2, (cuuint64_t)ss_h * 2, (cuuint64_t)ss_n * 2};
        // stride 2: one 9 x 9 pixel parity plane per load = every other pixel / row of an 18 x 18 window (conv_wgrad_halo.cu)
        cuuint32_t box[4] = {64, (cuuint32_t)(stride == 1 ? Geo<1>::PW : 2 * Geo<2>::PW), (cuuint32_t)(stride == 1 ? Geo<1>::PH : 2 * Geo<2>::PH), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = encode(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            gt_set_error("gt_conv2d_wgrad_f16 (wide halo): S tensor map rejected (CUresult %d)", (int)r);
            return -1;
        }
    }
    WWParams p;
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
    p.s_tiles = pl.s_tiles;
    p.nA = pl.nA;
    p.nB = pl.nB;
    p.wsA = workspace;
    p.wsB = workspace + (long long)pl.nA * 6 * UC * SC;
    for (int i = 0; i < 9; i++) pos_of[stride == 1 ? Geo<1>::tap_of(i) : Geo<2>::tap_of(i)] = i;
    *nA = pl.nA;
    *nB = pl.nB;
    return stride == 1 ? launch_wide<1>(tmU, tmS, p, pl, stream) : launch_wide<2>(tmU, tmS, p, pl, stream);
}
