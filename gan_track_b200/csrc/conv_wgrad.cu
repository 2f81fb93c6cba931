// conv_wgrad.cu -- weight gradient of the fp16 NHWC convolutions as a split-K implicit GEMM on tcgen05 + TMEM + TMA.
//
// Replaces cuDNN's wgrad, which the reference reaches through autograd of torch.nn.functional.conv2d /
// conv_transpose2d (OPS/conv2d_gradfix.py:37-45 on torch >= 1.11; the explicit formulation is in the dead-code
// Function at :150-201).
//
// Both convolution flavours reduce to one form.  With U the operand that is walked pixel by pixel and S the operand
// that is read at shifted (and possibly strided) positions,
//     G[u_ch][s_ch][r][s] = sum_{n,i,j} U[n,i,j,u_ch] * S[n, i*stride + r - pad, j*stride + s - pad, s_ch]
//     conv2d:            U = dy, S = x   -> dW[co][ci][r][s] = G          (weight layout [Cout,Cin,KH,KW])
//     conv_transpose2d:  U = x,  S = dy  -> dW[ci][co][r][s] = G          (weight layout [Cin,Cout,KH,KW])
// so dW[dim0][dim1] = G[u][s] in both.
//
// GEMM view: M = 128 U-channels, N = 64 or 128 S-channels, K = pixels.  The activation tiles arrive by TMA as
// [64 pixels][64 channels] boxes (128-byte rows, SWIZZLE_128B) -- for this GEMM that is an MN-major operand (the M / N
// index is contiguous, K walks rows), which tcgen05.mma reads through an MN-major shared-memory descriptor
// (LBO = distance between 64-channel boxes, SBO = 1024 B between 8-pixel groups).  One CTA owns up to three taps
// (one kernel row): it loads the U tile once per 64-pixel block and one shifted S tile per tap, and keeps one fp32
// accumulator per tap in TMEM (3 x 128 columns).  K is split across CTAs (grid.x); partial sums go to an fp32
// workspace and a second kernel reduces them in a fixed order (deterministic: replicas stay bit-identical, which
// the reference checks with misc.check_ddp_consistency) and writes fp16 into the weight layout.
#include "gt_common.cuh"
#include "gt_sm100.cuh"

using namespace sm100;

// conv_wgrad_halo.cu: halo-staged, tap-paired kernel for the stride-1 3x3 layers
bool gt_wgrad_halo_applicable(int N, int UH, int UW, int UC, int SC, int SH, int SW, int KH, int KW, int stride, int pad);
long long gt_wgrad_halo_workspace(int N, int UH, int UW, int UC, int SC, int stride);
int gt_launch_wgrad_halo(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s, long long ss_n, long long ss_h,
                         long long ss_w, int SH, int SW, int SC, int N, int stride, int pad, float* workspace, long long workspace_floats,
                         cudaStream_t stream);
void gt_wgrad_halo_enable_s2(int on);
// conv_wgrad_halo_wide.cu: N = 128 U channels per CTA, kernel rows split over two CTA types (stride-1 3x3 layers with UC % 128 == 0)
bool gt_wgrad_halo_wide_applicable(int N, int UH, int UW, int UC, int SC, int SH, int SW, int KH, int KW, int stride, int pad);
long long gt_wgrad_halo_wide_workspace(int N, int UH, int UW, int UC, int SC, int stride);
int gt_launch_wgrad_halo_wide(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s, long long ss_n,
                              long long ss_h, long long ss_w, int SH, int SW, int SC, int N, int stride, int pad, float* workspace,
                              long long workspace_floats, int* nA, int* nB, int* pos_of, cudaStream_t stream);
void gt_wgrad_halo_wide_enable(int on);
static int g_wgrad_variant = 0;   // 0 = auto (halo kernels where they apply), 1 = per-tap-row kernel only, 2 = halo kernel for stride 1 only,
                                  // 3 = no wide (N = 128) halo kernel
// > 0: at most this many pixels per split-K slice.  Set (per calling thread) around the fp16x3 entry points: the tensor core's truncating
// accumulation must stay below ~100 main-term updates per accumulator for fp32 accuracy (csrc/conv_f16x3.cu).
thread_local int t_wgrad_px_limit = 0;
extern "C" int gt_conv_wgrad_config(int variant) {
    const int old = g_wgrad_variant;
    g_wgrad_variant = variant;
    gt_wgrad_halo_enable_s2(variant != 2);
    gt_wgrad_halo_wide_enable(variant != 3);
    return old;
}

namespace {

constexpr int WM = 128;        // U channels per tile (UMMA M)
constexpr int WK = 64;         // pixels per k-block
constexpr int MAX_TG = 3;      // taps per CTA
constexpr int NTHREADS = 192;

struct WgradParams {
    int N, UH, UW;             // pixel domain of U
    int UC, SC;
    int stride, pad, KH, KW, ntaps;
    int bw_log2, bh_log2;      // 64-pixel box bw x bh x bn
    int tiles_w, tiles_h, tiles_n, num_tiles;
    int splits, s_tiles, tap_groups;
    float* ws;                 // [splits][ntaps][UC][SC]
};

template <int BN, int STAGES>
struct WSmem {
    static constexpr uint32_t BOX = WK * 128;                        // one [64 px][64 ch] box
    static constexpr uint32_t U_BYTES = 2 * BOX;                     // 128 U channels
    static constexpr uint32_t S_TAP_BYTES = (BN / 64) * BOX;         // BN S channels, one tap
    static constexpr uint32_t STAGE = U_BYTES + MAX_TG * S_TAP_BYTES;
    static constexpr uint32_t TILES = STAGES * STAGE;
    static constexpr uint32_t TOTAL = TILES + (2 * STAGES + 1) * 8 + 8 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS) conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmS,
                                                              const WgradParams p) {
    typedef WSmem<BN, STAGES> L;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + L::TILES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = full + 2 * STAGES;
    uint32_t* tmem_slot = (uint32_t*)(full + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int ut = blockIdx.y / p.s_tiles, st = blockIdx.y % p.s_tiles;
    const int tap0 = blockIdx.z * MAX_TG;
    const int ntap = min(MAX_TG, p.ntaps - tap0);

    constexpr uint32_t TMEM_COLS = (MAX_TG * BN) <= 256 ? 256 : 512;
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

    const int bw_log2 = p.bw_log2, bh_log2 = p.bh_log2, bn_log2 = 6 - bw_log2 - bh_log2;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t bytes = L::U_BYTES + (uint32_t)ntap * L::S_TAP_BYTES;
            for (int t = split; t < p.num_tiles; t += p.splits) {
                const int twi = t % p.tiles_w;
                const int rest = t / p.tiles_w;
                const int thi = rest % p.tiles_h, tni = rest / p.tiles_h;
                const int j0 = twi << bw_log2, i0 = thi << bh_log2, n0 = tni << bn_log2;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], bytes);
                uint8_t* sU = smem + stage * L::STAGE;
                tma_load_4d(sU, &tmU, &full[stage], ut * WM, j0, i0, n0);
                tma_load_4d(sU + L::BOX, &tmU, &full[stage], ut * WM + 64, j0, i0, n0);   // channels past UC read as zero
                for (int tp = 0; tp < ntap; tp++) {
                    const int tap = tap0 + tp;
                    const int r = tap / p.KW, s = tap - r * p.KW;
                    uint8_t* sS = sU + L::U_BYTES + tp * L::S_TAP_BYTES;
#pragma unroll
                    for (int h = 0; h < BN / 64; h++)
                        tma_load_4d(sS + h * L::BOX, &tmS, &full[stage], st * BN + h * 64, j0 * p.stride + s - p.pad, i0 * p.stride + r - p.pad, n0);
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
            constexpr uint32_t idesc = umma_idesc(WM, BN, 0, 1, 1);   // both operands MN-major
            int stage = 0;
            uint32_t phase = 0;
            bool first = true;
            for (int t = split; t < p.num_tiles; t += p.splits) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t u0 = smem_u32(smem + stage * L::STAGE);
                for (int tp = 0; tp < ntap; tp++) {
                    const uint32_t s0 = u0 + L::U_BYTES + tp * L::S_TAP_BYTES;
#pragma unroll
                    for (int k = 0; k < WK / 16; k++)
                        umma_f16(tmem_base + tp * BN, umma_smem_desc(u0 + k * 2048, L::BOX, 1024), umma_smem_desc(s0 + k * 2048, L::BOX, 1024), idesc,
                                 (uint32_t)(!first || k != 0));
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
        const int q = warp & 3;
        const int u = ut * WM + q * 32 + lane;
        const bool valid = u < p.UC;
        mbar_wait(tfull, 0);
        tc_fence_after();
        for (int tp = 0; tp < ntap; tp++) {
            float* wp = p.ws + (((long long)split * p.ntaps + tap0 + tp) * p.UC + u) * p.SC + st * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * BN + c * 32), r);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int v = 0; v < 8; v++)
                        *reinterpret_cast<uint4*>(wp + c * 32 + v * 4) = make_uint4(r[v * 4], r[v * 4 + 1], r[v * 4 + 2], r[v * 4 + 3]);
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

// dw[u][s][r][c] = sum over splits (in split order) of ws[split][tap][u][s]
// F32: fp16x3 route of the fp32 layers -- fp32 output, rescaled by the operands' power-of-two scales, only the first UCr x SCr channels
// (the operands are zero padded to multiples of 64 channels)
struct WgradOut {
    void* dw;
    long long ds_u, ds_s, ds_r, ds_c;
    int f32, UCr, SCr;
    const uint32_t *amax_u, *amax_s;
};
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int ntaps, int UC, int SC, int KW, const WgradOut o) {
    // one thread = one (tap, u, s) element: reads coalesced along s for every split (four splits in flight), 32-bit index math.
    // ~20 MB of partial sums whatever the layer (many splits x few channels ... few splits x many channels), so the thread count must
    // not depend on the channel count alone.
    const uint32_t pairs = (uint32_t)UC * (uint32_t)SC;
    const uint32_t per = (uint32_t)ntaps * pairs;
    const float inv = o.f32 ? (1.f / gt_scale_from_amax_bits(*o.amax_u)) * (1.f / gt_scale_from_amax_bits(*o.amax_s)) : 1.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        const uint32_t tap = i / pairs, pr = i - tap * pairs;
        const uint32_t u = pr / (uint32_t)SC, s = pr - u * (uint32_t)SC;
        if ((int)u >= o.UCr || (int)s >= o.SCr) continue;
        const float* p = ws + i;
        float acc = 0.f;
        int sp = 0;
        for (; sp + 4 <= splits; sp += 4) {                       // fixed order: ((((acc + p0) + p1) + p2) + p3) -- deterministic
            const float a0 = p[(size_t)sp * per], a1 = p[(size_t)(sp + 1) * per], a2 = p[(size_t)(sp + 2) * per], a3 = p[(size_t)(sp + 3) * per];
            acc = (((acc + a0) + a1) + a2) + a3;
        }
        for (; sp < splits; sp++) acc += p[(size_t)sp * per];
        const int r = (int)tap / KW, c = (int)tap - r * KW;
        const long long off = (long long)u * o.ds_u + (long long)s * o.ds_s + r * o.ds_r + c * o.ds_c;
        if (o.f32) ((float*)o.dw)[off] = acc * inv;
        else ((__half*)o.dw)[off] = __float2half_rn(acc);
    }
}

// 3x3 result of conv_wgrad_halo_wide.cu: tap positions 0..5 from region A ([nA][6][UC][SC]), 6..8 from region B ([nB][3][UC][SC]), each summed in
// slice order (deterministic); pos.p[tap] = position of tap r * 3 + s
struct TapPos {
    int p[9];
};
__global__ void __launch_bounds__(256) wgrad_reduce_ab_kernel(const float* __restrict__ wsA, int nA, const float* __restrict__ wsB, int nB, int UC, int SC,
                                                              const TapPos pos, const WgradOut o) {
    const uint32_t pairs = (uint32_t)UC * (uint32_t)SC;
    const uint32_t total = 9u * pairs;
    const float inv = o.f32 ? (1.f / gt_scale_from_amax_bits(*o.amax_u)) * (1.f / gt_scale_from_amax_bits(*o.amax_s)) : 1.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t tap = i / pairs, pr = i - tap * pairs;
        const uint32_t u = pr / (uint32_t)SC, s = pr - u * (uint32_t)SC;
        if ((int)u >= o.UCr || (int)s >= o.SCr) continue;
        const int ps = pos.p[tap];
        const bool a = ps < 6;
        const float* p = a ? wsA + (size_t)ps * pairs + pr : wsB + (size_t)(ps - 6) * pairs + pr;
        const size_t per = (size_t)(a ? 6 : 3) * pairs;
        const int splits = a ? nA : nB;
        float acc = 0.f;
        int sp = 0;
        for (; sp + 4 <= splits; sp += 4) {
            const float a0 = p[(size_t)sp * per], a1 = p[(size_t)(sp + 1) * per], a2 = p[(size_t)(sp + 2) * per], a3 = p[(size_t)(sp + 3) * per];
            acc = (((acc + a0) + a1) + a2) + a3;
        }
        for (; sp < splits; sp++) acc += p[(size_t)sp * per];
        const int r = (int)tap / 3, c = (int)tap - r * 3;
        const long long off = (long long)u * o.ds_u + (long long)s * o.ds_s + r * o.ds_r + c * o.ds_c;
        if (o.f32) ((float*)o.dw)[off] = acc * inv;
        else ((__half*)o.dw)[off] = __float2half_rn(acc);
    }
}

int next_pow2_log2(int v) {
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

struct WgradPlan {
    int bw_log2, bh_log2, tiles_w, tiles_h, tiles_n, num_tiles, splits, u_tiles, s_tiles, tap_groups, BN;
};

WgradPlan make_plan(int N, int UH, int UW, int UC, int SC, int ntaps) {
    WgradPlan pl;
    pl.bw_log2 = next_pow2_log2(UW);
    if (pl.bw_log2 > 4) pl.bw_log2 = 4;
    pl.bh_log2 = next_pow2_log2(UH);
    if (pl.bh_log2 > 6 - pl.bw_log2) pl.bh_log2 = 6 - pl.bw_log2;
    const int bn_log2 = 6 - pl.bw_log2 - pl.bh_log2;
    const int bw = 1 << pl.bw_log2, bh = 1 << pl.bh_log2, bn = 1 << bn_log2;
    pl.tiles_w = (UW + bw - 1) / bw;
    pl.tiles_h = (UH + bh - 1) / bh;
    pl.tiles_n = (N + bn - 1) / bn;
    pl.num_tiles = pl.tiles_w * pl.tiles_h * pl.tiles_n;
    pl.BN = (SC % 128 == 0) ? 128 : 64;
    pl.u_tiles = (UC + WM - 1) / WM;
    pl.s_tiles = SC / pl.BN;
    pl.tap_groups = (ntaps + MAX_TG - 1) / MAX_TG;
    const int per_split = pl.u_tiles * pl.s_tiles * pl.tap_groups;
    int splits = (2 * gt_num_sms() + per_split - 1) / per_split;   // about two CTAs per SM
    if (t_wgrad_px_limit > 0) {
        const long long px = (long long)N * UH * UW;
        const int need = (int)((px + t_wgrad_px_limit - 1) / t_wgrad_px_limit);
        if (splits < need) splits = need;
    }
    if (splits > pl.num_tiles) splits = pl.num_tiles;
    if (splits < 1) splits = 1;
    pl.splits = splits;
    return pl;
}

template <int BN, int STAGES>
int launch_wgrad(const CUtensorMap& tmU, const CUtensorMap& tmS, const WgradParams& p, int u_tiles, cudaStream_t stream) {
    typedef WSmem<BN, STAGES> L;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
        if (e != cudaSuccess) {
            gt_set_error("gt_conv2d_wgrad_f16: cannot reserve %u bytes of shared memory: %s", L::TOTAL, cudaGetErrorString(e));
            return GT_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid((unsigned)p.splits, (unsigned)(u_tiles * p.s_tiles), (unsigned)p.tap_groups);
    conv_wgrad_kernel<BN, STAGES><<<grid, NTHREADS, L::TOTAL, stream>>>(tmU, tmS, p);
    GT_CUDA_LAUNCH_CHECK("gt_conv2d_wgrad_f16");
    return GT_OK;
}

}  // namespace

static long long wgrad_workspace(int N, int UH, int UW, int UC, int SC, int KH, int KW);
extern "C" long long gt_conv2d_wgrad_workspace(int N, int UH, int UW, int UC, int SC, int KH, int KW) { return wgrad_workspace(N, UH, UW, UC, SC, KH, KW); }
// workspace of gt_conv2d_wgrad_f16x3 (more split-K slices than the fp16 plan: see t_wgrad_px_limit)
constexpr int F16X3_PX_LIMIT = 4608;
extern "C" long long gt_conv2d_wgrad_f16x3_workspace(int N, int UH, int UW, int UC, int SC, int KH, int KW) {
    t_wgrad_px_limit = F16X3_PX_LIMIT;
    const long long n = wgrad_workspace(N, UH, UW, UC, SC, KH, KW);
    t_wgrad_px_limit = 0;
    return n;
}
static long long wgrad_workspace(int N, int UH, int UW, int UC, int SC, int KH, int KW) {
    if (N <= 0 || UH <= 0 || UW <= 0 || UC <= 0 || SC <= 0 || KH <= 0 || KW <= 0) return 0;
    WgradPlan pl = make_plan(N, UH, UW, UC, SC, KH * KW);
    long long need = (long long)pl.splits * KH * KW * UC * SC;
    if (KH == 3 && KW == 3 && UC % 64 == 0 && SC % 64 == 0) {           // the halo kernel may take the call: cover its plans too
        for (int stride = 1; stride <= 2; stride++) {
            const long long h = gt_wgrad_halo_workspace(N, UH, UW, UC, SC, stride);
            if (h > need) need = h;
        }
        if (UC % 128 == 0) {
            for (int stride = 1; stride <= 2; stride++) {
                const long long h = gt_wgrad_halo_wide_workspace(N, UH, UW, UC, SC, stride);
                if (h > need) need = h;
            }
        }
    }
    return need;
}

static int wgrad_impl(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s, long long ss_n, long long ss_h,
                      long long ss_w, int SH, int SW, int SC, int N, int KH, int KW, int stride, int pad, const WgradOut& wo, float* workspace,
                      long long workspace_floats, void* stream) {
    void* dw = wo.dw;
    GT_REQUIRE(u && s && dw && workspace, "gt_conv2d_wgrad_f16: null pointer");
    GT_REQUIRE(N > 0 && UH > 0 && UW > 0 && SH > 0 && SW > 0, "gt_conv2d_wgrad_f16: empty tensor");
    GT_REQUIRE(UC % 64 == 0 && SC % 64 == 0, "gt_conv2d_wgrad_f16: channel counts (%d, %d) must be multiples of 64", UC, SC);
    GT_REQUIRE(KH >= 1 && KW >= 1 && KH * KW <= 9, "gt_conv2d_wgrad_f16: kernel %dx%d not supported", KH, KW);
    GT_REQUIRE(stride == 1 || stride == 2, "gt_conv2d_wgrad_f16: stride %d not supported", stride);
    GT_REQUIRE(pad >= 0 && pad < 8, "gt_conv2d_wgrad_f16: pad %d not supported", pad);
    GT_REQUIRE(((uintptr_t)u & 15) == 0 && ((uintptr_t)s & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "gt_conv2d_wgrad_f16: pointers must be 16-byte aligned");
    GT_REQUIRE(us_w % 8 == 0 && us_h % 8 == 0 && us_n % 8 == 0 && ss_w % 8 == 0 && ss_h % 8 == 0 && ss_n % 8 == 0,
               "gt_conv2d_wgrad_f16: strides must be multiples of 8 elements");
    const int ntaps = KH * KW;
    if (g_wgrad_variant != 1 && gt_wgrad_halo_wide_applicable(N, UH, UW, UC, SC, SH, SW, KH, KW, stride, pad)) {
        int nA = 0, nB = 0;
        TapPos pos;
        if (gt_launch_wgrad_halo_wide(u, us_n, us_h, us_w, UH, UW, UC, s, ss_n, ss_h, ss_w, SH, SW, SC, N, stride, pad, workspace, workspace_floats, &nA, &nB,
                                      pos.p, (cudaStream_t)stream) != 0)
            return GT_ERR_CUDA;
        long long g = ((long long)ntaps * UC * SC + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        wgrad_reduce_ab_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(workspace, nA, workspace + (long long)nA * 6 * UC * SC, nB, UC, SC, pos, wo);
        GT_CUDA_LAUNCH_CHECK("gt_conv2d_wgrad_f16 (reduce)");
        return GT_OK;
    }
    if (g_wgrad_variant != 1 && gt_wgrad_halo_applicable(N, UH, UW, UC, SC, SH, SW, KH, KW, stride, pad)) {
        const int splits = gt_launch_wgrad_halo(u, us_n, us_h, us_w, UH, UW, UC, s, ss_n, ss_h, ss_w, SH, SW, SC, N, stride, pad, workspace, workspace_floats,
                                                (cudaStream_t)stream);
        if (splits <= 0) return GT_ERR_CUDA;
        long long g = ((long long)ntaps * UC * SC + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        wgrad_reduce_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(workspace, splits, ntaps, UC, SC, KW, wo);
        GT_CUDA_LAUNCH_CHECK("gt_conv2d_wgrad_f16 (reduce)");
        return GT_OK;
    }
    WgradPlan pl = make_plan(N, UH, UW, UC, SC, ntaps);
    GT_REQUIRE(workspace_floats >= (long long)pl.splits * ntaps * UC * SC, "gt_conv2d_wgrad_f16: workspace too small");

    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.N = N;
    p.UH = UH;
    p.UW = UW;
    p.UC = UC;
    p.SC = SC;
    p.stride = stride;
    p.pad = pad;
    p.KH = KH;
    p.KW = KW;
    p.ntaps = ntaps;
    p.bw_log2 = pl.bw_log2;
    p.bh_log2 = pl.bh_log2;
    p.tiles_w = pl.tiles_w;
    p.tiles_h = pl.tiles_h;
    p.tiles_n = pl.tiles_n;
    p.num_tiles = pl.num_tiles;
    p.splits = pl.splits;
    p.s_tiles = pl.s_tiles;
    p.tap_groups = pl.tap_groups;
    p.ws = workspace;

    gt_encode_tiled_fn encode = gt_get_encode_tiled();
    GT_REQUIRE(encode != nullptr, "gt_conv2d_wgrad_f16: cuTensorMapEncodeTiled is not available from this driver");
    const int bw = 1 << pl.bw_log2, bh = 1 << pl.bh_log2, bn = 64 >> (pl.bw_log2 + pl.bh_log2);
    CUtensorMap tmU, tmS;
    {
        cuuint64_t dims[4] = {(cuuint64_t)UC, (cuuint64_t)UW, (cuuint64_t)UH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)us_w * 2, (cuuint64_t)us_h * 2, (cuuint64_t)us_n * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmU, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(u), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_wgrad_f16: U tensor map rejected (CUresult %d)", (int)r);
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)SC, (cuuint64_t)SW, (cuuint64_t)SH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ss_w * 2, (cuuint64_t)ss_h * 2, (cuuint64_t)ss_n * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = encode(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_wgrad_f16: S tensor map rejected (CUresult %d)", (int)r);
    }
    cudaStream_t stm = (cudaStream_t)stream;
    int rc = (pl.BN == 128) ? launch_wgrad<128, 3>(tmU, tmS, p, pl.u_tiles, stm) : launch_wgrad<64, 4>(tmU, tmS, p, pl.u_tiles, stm);
    if (rc != GT_OK) return rc;
    const int block = 256;
    long long g = ((long long)ntaps * UC * SC + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    wgrad_reduce_kernel<<<(int)g, block, 0, stm>>>(workspace, pl.splits, ntaps, UC, SC, KW, wo);
    GT_CUDA_LAUNCH_CHECK("gt_conv2d_wgrad_f16 (reduce)");
    return GT_OK;
}

extern "C" int gt_conv2d_wgrad_f16(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s, long long ss_n,
                                   long long ss_h, long long ss_w, int SH, int SW, int SC, int N, int KH, int KW, int stride, int pad, void* dw,
                                   long long ds_u, long long ds_s, long long ds_r, long long ds_c, float* workspace, long long workspace_floats,
                                   void* stream) {
    WgradOut wo = {dw, ds_u, ds_s, ds_r, ds_c, 0, UC, SC, nullptr, nullptr};
    return wgrad_impl(u, us_n, us_h, us_w, UH, UW, UC, s, ss_n, ss_h, ss_w, SH, SW, SC, N, KH, KW, stride, pad, wo, workspace, workspace_floats, stream);
}

// Weight gradient of an fp32 layer on the fp16x3 route (csrc/conv_f16x3.cu): u / s are the batch-concatenated fp16 splits
// ([3N,H,W,UC] images (hi, hi, lo) and [3N,H,W,SC] images (hi, lo, hi), channel counts padded to multiples of 64), `N` counts all 3N
// images, dw is fp32 and receives the first UC_real x SC_real channels rescaled by the operands' scales.
extern "C" int gt_conv2d_wgrad_f16x3(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s, long long ss_n,
                                     long long ss_h, long long ss_w, int SH, int SW, int SC, int N, int KH, int KW, int stride, int pad, void* dw,
                                     long long ds_u, long long ds_s, long long ds_r, long long ds_c, int UC_real, int SC_real, const void* amax_u,
                                     const void* amax_s, float* workspace, long long workspace_floats, void* stream) {
    GT_REQUIRE(amax_u && amax_s, "gt_conv2d_wgrad_f16x3: null scale pointer");
    GT_REQUIRE(UC_real >= 1 && UC_real <= UC && SC_real >= 1 && SC_real <= SC, "gt_conv2d_wgrad_f16x3: bad channel counts");
    WgradOut wo = {dw, ds_u, ds_s, ds_r, ds_c, 1, UC_real, SC_real, (const uint32_t*)amax_u, (const uint32_t*)amax_s};
    t_wgrad_px_limit = F16X3_PX_LIMIT;
    const int rc = wgrad_impl(u, us_n, us_h, us_w, UH, UW, UC, s, ss_n, ss_h, ss_w, SH, SW, SC, N, KH, KW, stride, pad, wo, workspace, workspace_floats, stream);
    t_wgrad_px_limit = 0;
    return rc;
}
