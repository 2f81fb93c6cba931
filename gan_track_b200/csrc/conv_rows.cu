// conv_rows.cu -- row-streaming 3x3 stride-1 convolution for the 64 -> 64 channel layers at the top resolutions (G b256.conv1,
// D b256.conv0 and their data gradients: the kernel with the largest share of the training step), fp16 NHWC, tcgen05 + TMEM + TMA.
//
// Why a second formulation.  Measured on B200 (tools/umma_probe3.cu, profiles/r02_umma_probe.txt): an SS-mode tcgen05.mma of shape
// M128 x N x K16 takes max(N / 2, 32 + N / 4) clocks -- the two operands stream from shared memory at 128 B/clk together -- so an
// N = 64 implicit GEMM (pixels x output channels) cannot exceed 2/3 of the tensor-core rate, N >= 128 can reach it.  With only 64
// output channels the N dimension is widened with the KERNEL ROWS instead:
//
//     out[y][x][co] = sum_{dy,dx,ci} in[y + dy][x + dx][ci] * w[dy][dx][co][ci]
//
//   * A (M = 128) = a strip of 128 consecutive pixels of ONE input row r, shifted by dx pixels (start address + dx * 128 bytes);
//   * B (N = 192) = for one dx the three weight slabs [w(dy=+1) ; w(dy=0) ; w(dy=-1)] stacked, so one MMA adds input row r into the
//     THREE output rows r-1, r, r+1 at once; their accumulators are three adjacent 64-column blocks of TMEM;
//   * input rows stream top to bottom: every step (= one input row, 3 dx x 4 K-slices = 12 MMAs of N = 192, all columns useful)
//     completes one output row and opens a new one.  TMEM is a ring of eight 64-column blocks, the window (r-1, r, r+1) slides by one
//     block per step; where it wraps, or where the newly opened block must be overwritten instead of accumulated, an MMA is issued in
//     two or three pieces.
// Per step: one 16.6 KB TMA row load (contiguous in global memory), 12-13 MMAs (~1170 clk), one 16 KB TMA row store -- every input
// byte is fetched once per strip (+ 2 halo pixels per row), the weights (72 KB) stay in shared memory for the whole persistent CTA.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue (TMEM -> registers -> [bias_act] -> fp16 -> swizzled shared
// memory -> TMA store): four warps for the plain store; EIGHT with the fused bias_act (two per TMEM lane quarter, each taking 32 of the 64
// output channels, the bias in registers) -- with four, one warp per scheduler has nothing to hide the dependent conversion / activation
// chains behind and the kernel becomes epilogue-paced (measured: 282 us fused against 103 us for the plain kernel).  Work split: the N * strips * H output rows are cut into contiguous ranges, one per CTA.
#include "conv_common.cuh"

using namespace sm100;

namespace {

constexpr int NTHREADS = 192;                             // plain store: producer + issuer + 4 epilogue warps
constexpr int NTHREADS_EP = 320;                          // fused bias_act: 8 epilogue warps (two per TMEM lane quarter, 32 of the 64 columns each)
constexpr int STRIP = 128;                                // output pixels per step = UMMA M
constexpr int ROW_PX = STRIP + 2;                         // + one halo pixel each side
constexpr uint32_t ROW_BYTES = ROW_PX * 128;              // 16640: 130 pixels x 64 channels fp16
constexpr uint32_t ROW_SLOT = 17 * 1024;                  // slot pitch (1024-byte aligned for the 128-byte swizzle)
constexpr int SLOTS = 6;
constexpr uint32_t W_SLAB = 64 * 128;                     // one tap: 64 output channels x 64 input channels fp16
constexpr uint32_t W_BYTES = 9 * W_SLAB;
constexpr uint32_t STAGE_BYTES = STRIP * 128;             // one output row of the strip
constexpr int NSTAGE = 2;
constexpr int NBLK = 8;                                   // TMEM ring: 8 blocks of 64 fp32 columns
constexpr uint32_t SMEM_TILES = W_BYTES + SLOTS * ROW_SLOT + NSTAGE * STAGE_BYTES;
constexpr uint32_t NBARS = 1 + 2 * SLOTS + 2 * NBLK;
constexpr uint32_t SMEM_TOTAL = SMEM_TILES + NBARS * 8 + 16 + 1024;

struct RowsParams {
    int N, H, W, xs;                 // xs = strips per image row
    long long total_rows;            // N * xs * H output rows
    long long rows_per_cta;
    int slab_of[9];                  // weight slab (tap index in the packed weight tensor) for position dxi * 3 + blk (blk: dy = +1, 0, -1)
    // optional fused bias_act (CONV_F16_EP semantics of conv_common.cuh)
    const __half* ep_bias;
    int ep_act;
    float ep_alpha, ep_gain, ep_clamp;
};

__device__ __forceinline__ void tma_store_4d(const void* smem_src, const CUtensorMap* m, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)m), "r"(smem_u32(smem_src)), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

struct Seg {
    int n, x0, ya, yb;
};
__device__ __forceinline__ Seg segment_at(const RowsParams& p, long long row, long long row_end) {
    Seg s;
    const long long strip = row / p.H;
    s.ya = (int)(row - strip * p.H);
    const long long left = row_end - row;
    s.yb = (left < (long long)(p.H - s.ya)) ? s.ya + (int)left : p.H;
    s.n = (int)(strip / p.xs);
    s.x0 = (int)(strip - (long long)s.n * p.xs) * STRIP;
    return s;
}


// 32 accumulator columns of one pixel -> [bias_act] -> fp16 -> four swizzled 16-byte chunks of the pixel's staged 128-byte row.  One
// straight-line instantiation per (activation, clamp): 16 independent half2 chains for the scheduler to interleave.
template <bool EP, int ACT, bool CLAMP>
__device__ __forceinline__ void stage_cols(const uint32_t (&v)[32], const __half2 (&bias_r)[16], bool has_bias, const hot::Params& hp, uint32_t srow, int c,
                                           int m) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        __half2 h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            h[k] = __floats2half2_rn(__uint_as_float(v[g * 8 + 2 * k]), __uint_as_float(v[g * 8 + 2 * k + 1]));
            if (EP) {
                float2 u = __half22float2(h[k]);                    // round to fp16 first: the reference materialises the convolution output
                if (has_bias) u = __fadd2_rn(u, __half22float2(bias_r[g * 4 + k]));
                h[k] = __float22half2_rn(hot::fwd<ACT, CLAMP>(u, hp));
            }
        }
        uint4 o4;
        o4.x = *reinterpret_cast<uint32_t*>(&h[0]);
        o4.y = *reinterpret_cast<uint32_t*>(&h[1]);
        o4.z = *reinterpret_cast<uint32_t*>(&h[2]);
        o4.w = *reinterpret_cast<uint32_t*>(&h[3]);
        const uint32_t chunk = (uint32_t)(c * 4 + g);                           // 16-byte chunk of the pixel's 128-byte row
        sts128(srow + ((chunk ^ ((uint32_t)m & 7u)) << 4), o4);                 // SWIZZLE_128B: chunk ^ (row mod 8)
    }
}

template <bool EP>
__global__ void __launch_bounds__(EP ? NTHREADS_EP : NTHREADS, 1) conv_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmY, const RowsParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;
    uint8_t* sRows = smem + W_BYTES;
    uint8_t* sStage = sRows + SLOTS * ROW_SLOT;
    uint64_t* w_full = (uint64_t*)(smem + SMEM_TILES);
    uint64_t* row_full = w_full + 1;
    uint64_t* row_empty = row_full + SLOTS;
    uint64_t* acc_full = row_empty + SLOTS;
    uint64_t* acc_empty = acc_full + NBLK;
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + NBLK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmY);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(w_full, 1);
        for (int i = 0; i < SLOTS; i++) {
            mbar_init(&row_full[i], 1);
            mbar_init(&row_empty[i], 1);
        }
        for (int i = 0; i < NBLK; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], EP ? 8 : 4);          // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long row_begin = (long long)blockIdx.x * p.rows_per_cta;
    long long row_end = row_begin + p.rows_per_cta;
    if (row_end > p.total_rows) row_end = p.total_rows;

    if (warp == 0) {
        if (elect_one()) {
            // resident weights: nine 8 KB slabs, position dxi * 3 + blk
            mbar_arrive_expect_tx(w_full, W_BYTES);
#pragma unroll 1
            for (int i = 0; i < 9; i++) tma_load_3d(sW + i * W_SLAB, &tmB, w_full, 0, 0, p.slab_of[i]);
            uint32_t j = 0;
            for (long long row = row_begin; row < row_end;) {
                const Seg s = segment_at(p, row, row_end);
                for (int r = s.ya - 1; r <= s.yb; r++, j++) {
                    const uint32_t slot = j % SLOTS;
                    mbar_wait(&row_empty[slot], ((j / SLOTS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&row_full[slot], ROW_BYTES);
                    tma_load_4d(sRows + slot * ROW_SLOT, &tmA, &row_full[slot], 0, s.x0 - 1, r, s.n);      // x = -1, x >= W, y = -1, y = H read as zero
                }
                row += s.yb - s.ya;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t id64 = umma_idesc(128, 64, 0, 0, 0), id128 = umma_idesc(128, 128, 0, 0, 0), id192 = umma_idesc(128, 192, 0, 0, 0);
            const uint64_t dB = umma_smem_desc(smem_u32(sW), 0, 1024);
            const uint64_t dA0 = umma_smem_desc(smem_u32(sRows), 0, 1024);
            mbar_wait(w_full, 0);
            tc_fence_after();
            uint32_t j = 0;
            for (long long row = row_begin; row < row_end;) {
                const Seg s = segment_at(p, row, row_end);
                for (int r = s.ya - 1; r <= s.yb; r++, j++) {
                    const uint32_t slot = j % SLOTS;
                    const uint32_t b0 = j % NBLK;                        // window = blocks b0, b0+1, b0+2 (mod 8) = output rows r-1, r, r+1
                    const uint32_t fresh = (j + 2) % NBLK;               // block of output row r+1: first contribution -> overwrite
                    if (j + 2 >= NBLK) {
                        mbar_wait(&acc_empty[fresh], (((j + 2) / NBLK) - 1) & 1);
                    }
                    mbar_wait(&row_full[slot], (j / SLOTS) & 1);
                    tc_fence_after();
                    const uint64_t dA = dA0 + (uint64_t)((slot * ROW_SLOT) >> 4);
                    const uint32_t c0 = tmem_base + b0 * 64;
                    if (b0 <= 5) {
#pragma unroll
                        for (int q = 0; q < 12; q++) {
                            const int dx = q / 4, k = q % 4;
                            const uint64_t a = dA + (uint64_t)((dx * 128 + k * 32) >> 4);
                            const uint64_t b = dB + (uint64_t)((dx * 3 * (int)W_SLAB + k * 32) >> 4);
                            if (q == 0) {
                                umma_f16(c0, a, b, id128, 1u);
                                umma_f16(c0 + 128, a, b + (uint64_t)((2 * W_SLAB) >> 4), id64, 0u);
                            } else {
                                umma_f16(c0, a, b, id192, 1u);
                            }
                        }
                    } else if (b0 == 6) {                                 // blocks 6, 7 | 0
#pragma unroll
                        for (int q = 0; q < 12; q++) {
                            const int dx = q / 4, k = q % 4;
                            const uint64_t a = dA + (uint64_t)((dx * 128 + k * 32) >> 4);
                            const uint64_t b = dB + (uint64_t)((dx * 3 * (int)W_SLAB + k * 32) >> 4);
                            umma_f16(c0, a, b, id128, 1u);
                            umma_f16(tmem_base, a, b + (uint64_t)((2 * W_SLAB) >> 4), id64, q == 0 ? 0u : 1u);
                        }
                    } else {                                              // block 7 | 0, 1
#pragma unroll
                        for (int q = 0; q < 12; q++) {
                            const int dx = q / 4, k = q % 4;
                            const uint64_t a = dA + (uint64_t)((dx * 128 + k * 32) >> 4);
                            const uint64_t b = dB + (uint64_t)((dx * 3 * (int)W_SLAB + k * 32) >> 4);
                            umma_f16(c0, a, b, id64, 1u);
                            if (q == 0) {
                                umma_f16(tmem_base, a, b + (uint64_t)(W_SLAB >> 4), id64, 1u);
                                umma_f16(tmem_base + 64, a, b + (uint64_t)((2 * W_SLAB) >> 4), id64, 0u);
                            } else {
                                umma_f16(tmem_base, a, b + (uint64_t)(W_SLAB >> 4), id128, 1u);
                            }
                        }
                    }
                    umma_commit(&row_empty[slot]);          // the input row may be overwritten once these MMAs have read it
                    umma_commit(&acc_full[b0]);             // output row r-1 (block b0) is complete
                }
                row += s.yb - s.ya;
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int m = q * 32 + lane;                                     // pixel of the strip = TMEM lane
        const bool issuer = (warp == 2 && lane == 0);
        constexpr int EP_THREADS = EP ? 256 : 128;
        const int half = EP ? ((warp - 2) >> 2) : 0;                     // fused epilogue: which 32 of the 64 output channels this warp takes
        const hot::Params hp = hot::make_params(p.ep_alpha, p.ep_gain, p.ep_clamp);
        const bool has_bias = EP && p.ep_bias != nullptr;
        const int ep_mode = EP ? ((p.ep_act == hot::LRELU ? 2 : 0) + (p.ep_clamp >= 0.f ? 1 : 0)) : 0;
        __half2 bias_r[16];                                              // the warp's 32 bias values
        if (EP) {
#pragma unroll
            for (int g = 0; g < 4; g++) {
                Vec16<__half> bv;
                if (has_bias) bv = ld16(p.ep_bias + half * 32 + g * 8);
#pragma unroll
                for (int k = 0; k < 4; k++) bias_r[g * 4 + k] = has_bias ? reinterpret_cast<const __half2*>(bv.v)[k] : __floats2half2_rn(0.f, 0.f);
            }
        }
        uint32_t j = 0, nstore = 0;
        for (long long row = row_begin; row < row_end;) {
            const Seg s = segment_at(p, row, row_end);
            for (int r = s.ya - 1; r <= s.yb; r++, j++) {
                const uint32_t b0 = j % NBLK;
                const int y = r - 1;
                const bool valid = y >= s.ya && y < s.yb;
                mbar_wait(&acc_full[b0], (j / NBLK) & 1);
                tc_fence_after();
                if (valid) {
                    uint8_t* stage = sStage + (nstore % NSTAGE) * STAGE_BYTES;
                    if (issuer) bulk_wait_read<NSTAGE - 1>();            // the store that last read this staging buffer has drained it
                    named_bar_sync(1, EP_THREADS);
                    const uint32_t srow = smem_u32(stage) + (uint32_t)m * 128u;
#pragma unroll
                    for (int cc = 0; cc < (EP ? 1 : 2); cc++) {
                        const int c = EP ? half : cc;
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + b0 * 64 + (uint32_t)(c * 32), v);
                        tmem_ld_wait();
                        if (!EP) stage_cols<false, hot::LINEAR, false>(v, bias_r, false, hp, srow, c, m);
                        else if (ep_mode == 3) stage_cols<true, hot::LRELU, true>(v, bias_r, has_bias, hp, srow, c, m);
                        else if (ep_mode == 2) stage_cols<true, hot::LRELU, false>(v, bias_r, has_bias, hp, srow, c, m);
                        else if (ep_mode == 1) stage_cols<true, hot::LINEAR, true>(v, bias_r, has_bias, hp, srow, c, m);
                        else stage_cols<true, hot::LINEAR, false>(v, bias_r, has_bias, hp, srow, c, m);
                    }
                    tc_fence_before();
                    fence_proxy_async();
                    named_bar_sync(1, EP_THREADS);
                    if (issuer) {
                        tma_store_4d(stage, &tmY, 0, s.x0, y, s.n);      // pixels past the image width are clipped by the TMA unit
                        bulk_commit();
                    }
                    nstore++;
                } else {
                    tc_fence_before();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[b0]);
            }
            row += s.yb - s.ya;
        }
        if (issuer) bulk_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

static int g_conv_rows_enabled = 1;
extern "C" int gt_conv_rows_config(int enabled) {
    const int old = g_conv_rows_enabled;
    g_conv_rows_enabled = enabled;
    return old;
}

bool gt_conv_rows_applicable(const ConvParams& p, int H, int W) {
    if (!g_conv_rows_enabled || p.Cin != 64 || p.Cout != 64 || p.nphases != 1 || p.in_stride != 1 || p.out_stride != 1) return false;
    const int mode = conv_mode(p);
    if (mode != CONV_F16 && mode != CONV_F16_EP) return false;
    if (p.ep_add) return false;                  // the residual form of the epilogue lives in conv_store32 (halo / per-tap kernels)
    const ConvPhase& ph = p.ph[0];
    if (ph.ntaps != 9 || ph.OHp != H || ph.OWp != W || ph.y_off != 0) return false;
    bool seen[9] = {false, false, false, false, false, false, false, false, false};
    for (int t = 0; t < 9; t++) {
        if (ph.tdy[t] < -1 || ph.tdy[t] > 1 || ph.tdx[t] < -1 || ph.tdx[t] > 1) return false;
        seen[(ph.tdx[t] + 1) * 3 + (1 - ph.tdy[t])] = true;
    }
    for (int i = 0; i < 9; i++)
        if (!seen[i]) return false;
    return W >= STRIP && H >= 8;
}

int gt_launch_conv_rows(const void* x, long long xs_n, long long xs_h, long long xs_w, int H, int W, const void* wpacked, int ntaps_total, const ConvParams& p,
                        cudaStream_t stream) {
    RowsParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.N = p.N;
    rp.H = H;
    rp.W = W;
    rp.xs = (W + STRIP - 1) / STRIP;
    rp.total_rows = (long long)p.N * rp.xs * H;
    int grid = gt_num_sms();
    if ((long long)grid > rp.total_rows / 8) grid = (int)(rp.total_rows / 8 > 0 ? rp.total_rows / 8 : 1);
    rp.rows_per_cta = (rp.total_rows + grid - 1) / grid;
    grid = (int)((rp.total_rows + rp.rows_per_cta - 1) / rp.rows_per_cta);
    const ConvPhase& ph = p.ph[0];
    for (int t = 0; t < 9; t++) rp.slab_of[(ph.tdx[t] + 1) * 3 + (1 - ph.tdy[t])] = ph.tw[t];
    rp.ep_bias = p.ep_bias;
    rp.ep_act = p.ep_act;
    rp.ep_alpha = p.ep_alpha;
    rp.ep_gain = p.ep_gain;
    rp.ep_clamp = p.ep_clamp;

    gt_encode_tiled_fn encode = gt_get_encode_tiled();
    GT_REQUIRE(encode != nullptr, "gt_conv2d_igemm_f16 (rows): cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tmA, tmB, tmY;
    {
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)xs_w * 2, (cuuint64_t)xs_h * 2, (cuuint64_t)xs_n * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)ROW_PX, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm_f16 (rows): activation tensor map rejected (CUresult %d)", (int)r);
    }
    {
        cuuint64_t dims[3] = {64, 64, (cuuint64_t)ntaps_total};
        cuuint64_t strides[2] = {64 * 2, 64 * 64 * 2};
        cuuint32_t box[3] = {64, 64, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(wpacked), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm_f16 (rows): weight tensor map rejected (CUresult %d)", (int)r);
    }
    {
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)p.ys_w * 2, (cuuint64_t)p.ys_h * 2, (cuuint64_t)p.ys_n * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)STRIP, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, p.y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm_f16 (rows): output tensor map rejected (CUresult %d)", (int)r);
    }
    static bool configured[2] = {false, false};
    const int ep = conv_mode(p) == CONV_F16_EP ? 1 : 0;
    if (!configured[ep]) {
        cudaError_t e = ep ? cudaFuncSetAttribute(conv_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL)
                           : cudaFuncSetAttribute(conv_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL);
        if (e != cudaSuccess) {
            gt_set_error("gt_conv2d_igemm_f16 (rows): cannot reserve %u bytes of shared memory: %s", SMEM_TOTAL, cudaGetErrorString(e));
            return GT_ERR_CUDA;
        }
        configured[ep] = true;
    }
    if (ep) conv_rows_kernel<true><<<grid, NTHREADS_EP, SMEM_TOTAL, stream>>>(tmA, tmB, tmY, rp);
    else conv_rows_kernel<false><<<grid, NTHREADS, SMEM_TOTAL, stream>>>(tmA, tmB, tmY, rp);
    GT_CUDA_LAUNCH_CHECK("gt_conv2d_igemm_f16 (rows)");
    return GT_OK;
}
