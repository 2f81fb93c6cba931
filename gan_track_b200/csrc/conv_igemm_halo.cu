// conv_igemm_halo.cu -- persistent, halo-staged variant of the fp16 NHWC implicit-GEMM convolution for every
// stride-1 phase (3x3 / 1x1 convolutions, transposed stride-1, and the four parity phases of the transposed stride-2
// up-convolution).  Same contract and call sites as conv_igemm.cu; this kernel exists because the per-tap kernel is
// bound by L2 -> SM traffic on the 64/128-channel layers (ncu, profiles/r01_ncu_full_summary.txt: 3.6 GB through the
// crossbar for 0.54 GB of algorithmic traffic -- every tap re-read the activation tile, every 128-pixel tile re-read
// the weights).
//
// What changes:
//   * CTA tile = MT sub-tiles of 8 x 16 output pixels side by side (MT*BN = 256 accumulator columns; MT = 2 with BN = 128,
//     MT = 4 with BN = 64).  ONE TMA box per 64-channel chunk brings the whole input footprint of the tile including
//     the halo ((8*MT + ext_x) x (16 + ext_y) pixels); all taps and all sub-tiles read it in place: a tap is just a
//     different start address in the UMMA shared-memory descriptor, the 8-pixel tile rows are the 8-row core
//     matrices and the halo row pitch is the descriptor's stride byte offset.  (The 128-byte swizzle is a function
//     of the shared-memory address bits, so a tile row may start at any 128-byte row of the staged box.)
//   * The weight tile of a tap is loaded once per chunk and used by all MT sub-tiles.
//   * Persistent CTAs (grid = #SMs) walk the tile list; TMEM holds two accumulator sets (2 x 256 columns) so the
//     epilogue of tile i overlaps the MMAs of tile i+1.
// Pipelines: A-halo buffers (2) and weight-tap ring (SB stages), both full/empty mbarrier pairs; TMEM full/empty pair.
#include "conv_common.cuh"

using namespace sm100;

namespace {

constexpr int NTHREADS = 192;
constexpr int NTHREADS_EP = 320;     // fused bias_act (+ residual): eight epilogue warps, as in conv_igemm_halo2.cu / conv_rows.cu
constexpr int SUB_W = 8, SUB_H = 16;   // one UMMA M=128 sub-tile: 16 rows of 8 pixels

template <int BN, int MT, int SB, int NBUF>
struct HaloSmem {
    static constexpr uint32_t A_MAX_PIX = (SUB_W * MT + 2) * (SUB_H + 2);
    static constexpr uint32_t A_BYTES = ((A_MAX_PIX * 128 + 1023) / 1024) * 1024;
    static constexpr uint32_t B_BYTES = BN * 128;
    static constexpr uint32_t TILES = 2 * A_BYTES + SB * B_BYTES;
    static constexpr uint32_t NBARS = 4 + 2 * SB + 2 * NBUF;
    static constexpr uint32_t TOTAL = TILES + NBARS * 8 + 16 + 1024;
};

struct Item {
    int phase, nt, ox0, oy0, n;
    bool valid;
};

__device__ __forceinline__ Item decode_item(const ConvParams& p, int item, int tile_w) {
    Item it;
    it.nt = item % p.n_tiles;
    int t = item / p.n_tiles;
    const int twi = t % p.tiles_w;
    t /= p.tiles_w;
    const int thi = t % p.tiles_h;
    t /= p.tiles_h;
    it.n = t % p.N;
    it.phase = t / p.N;
    it.ox0 = twi * tile_w;
    it.oy0 = thi * SUB_H;
    it.valid = it.ox0 < p.ph[it.phase].OWp && it.oy0 < p.ph[it.phase].OHp;
    return it;
}

template <int BN, int MT, int SB, int NBUF, int MODE>
__global__ void __launch_bounds__(MODE == CONV_F16_EP ? NTHREADS_EP : NTHREADS, 1) conv_igemm_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                       const ConvParams p, const int total_items) {
    typedef HaloSmem<BN, MT, SB, NBUF> L;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + 2 * L::A_BYTES;
    uint64_t* a_full = (uint64_t*)(smem + L::TILES);
    uint64_t* a_empty = a_full + 2;
    uint64_t* b_full = a_full + 4;
    uint64_t* b_empty = b_full + SB;
    uint64_t* t_full = b_empty + SB;
    uint64_t* t_empty = t_full + NBUF;
    uint32_t* tmem_slot = (uint32_t*)(t_empty + NBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int TILE_W = SUB_W * MT;
    constexpr uint32_t ACC_COLS = MT * BN;
    constexpr uint32_t TMEM_COLS = NBUF * ACC_COLS;
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < NBUF; i++) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], MODE == CONV_F16_EP ? 8 : 4);   // one arrival per epilogue warp
        }
        for (int i = 0; i < SB; i++) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr int kel = 64;                                    // channels per 128-byte chunk
    const int kchunks = p.Cin / kel;
    const uint32_t a_bytes = (uint32_t)(p.halo_w * p.halo_h) * 128u;
    const uint32_t row_pitch = (uint32_t)p.halo_w * 128u;

    if (warp == 0) {
        if (elect_one()) {
            uint32_t a_it = 0, b_it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const Item it = decode_item(p, item, TILE_W);
                if (!it.valid) continue;
                const ConvPhase& ph = p.ph[it.phase];
                const int cx = it.ox0 + p.dx_min[it.phase], cy = it.oy0 + p.dy_min[it.phase];
                for (int kc = 0; kc < kchunks; kc++) {
                    const uint32_t ab = a_it & 1;
                    mbar_wait(&a_empty[ab], ((a_it >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&a_full[ab], a_bytes);
                    tma_load_4d(sA + ab * L::A_BYTES, &tmA, &a_full[ab], kc * kel, cx, cy, it.n);
                    a_it++;
                    for (int t = 0; t < ph.ntaps; t++) {
                        const uint32_t bs = b_it % SB;
                        mbar_wait(&b_empty[bs], ((b_it / SB) & 1) ^ 1);
                        mbar_arrive_expect_tx(&b_full[bs], L::B_BYTES);
                        tma_load_3d(sB + bs * L::B_BYTES, &tmB, &b_full[bs], kc * kel, it.nt * BN, ph.tw[t]);
                        b_it++;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(128, BN, 0, 0, 0);
            uint32_t a_it = 0, b_it = 0, t_it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const Item it = decode_item(p, item, TILE_W);
                if (!it.valid) continue;
                const ConvPhase& ph = p.ph[it.phase];
                const int dy0 = p.dy_min[it.phase], dx0 = p.dx_min[it.phase];
                const uint32_t buf = t_it % NBUF;
                mbar_wait(&t_empty[buf], ((t_it / NBUF) & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * ACC_COLS;
                for (int kc = 0; kc < kchunks; kc++) {
                    const uint32_t ab = a_it & 1;
                    mbar_wait(&a_full[ab], (a_it >> 1) & 1);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + ab * L::A_BYTES);
                    for (int t = 0; t < ph.ntaps; t++) {
                        const uint32_t bs = b_it % SB;
                        mbar_wait(&b_full[bs], (b_it / SB) & 1);
                        tc_fence_after();
                        const uint32_t b0 = smem_u32(sB + bs * L::B_BYTES);
                        const uint32_t at = a0 + (uint32_t)(ph.tdy[t] - dy0) * row_pitch + (uint32_t)(ph.tdx[t] - dx0) * 128u;
#pragma unroll
                        for (int j = 0; j < MT; j++) {
#pragma unroll
                            for (int k = 0; k < 4; k++)        // 32 bytes of every row per MMA: K = 16 fp16
                                umma_f16(acc + j * BN, umma_smem_desc(at + j * (SUB_W * 128) + k * 32, 0, row_pitch),
                                         umma_smem_desc(b0 + k * 32, 0, 1024), idesc, (uint32_t)((kc | t | k) != 0));
                        }
                        umma_commit(&b_empty[bs]);
                        b_it++;
                    }
                    umma_commit(&a_empty[ab]);
                    a_it++;
                }
                umma_commit(&t_full[buf]);
                t_it++;
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        constexpr int CSETS = MODE == CONV_F16_EP ? 2 : 1;            // warp sets sharing a lane quarter take alternating column chunks
        const int cset = MODE == CONV_F16_EP ? ((warp - 2) >> 2) : 0;
        const int m = q * 32 + lane;
        const int lw = m & (SUB_W - 1), lh = m >> 3;
        uint32_t t_it = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item it = decode_item(p, item, TILE_W);
            if (!it.valid) continue;
            const ConvPhase& ph = p.ph[it.phase];
            const uint32_t buf = t_it % NBUF;
            mbar_wait(&t_full[buf], (t_it / NBUF) & 1);
            tc_fence_after();
            const int a = it.oy0 + lh;
#pragma unroll 1
            for (int j = 0; j < MT; j++) {
                const int b = it.ox0 + j * SUB_W + lw;
                const bool valid = (a < ph.OHp) && (b < ph.OWp) && !p.dbg_no_store;
                const long long yoff = (long long)it.n * p.ys_n + (long long)(a * p.out_stride + ph.off_y) * p.ys_h +
                                       (long long)(b * p.out_stride + ph.off_x) * p.ys_w + it.nt * BN + ph.y_off;
#pragma unroll 1
                for (int c = cset; c < BN / 32; c += CSETS) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + (uint32_t)(j * BN + c * 32), r);
                    tmem_ld_wait();
                    if (valid) conv_store32<MODE>(p, yoff + c * 32, it.nt * BN + c * 32, r);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
            t_it++;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BN, int MT, int SB, int NBUF, int MODE>
int launch_halo_m(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int total_items, cudaStream_t stream) {
    typedef HaloSmem<BN, MT, SB, NBUF> L;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_halo_kernel<BN, MT, SB, NBUF, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
        if (e != cudaSuccess) {
            gt_set_error("gt_conv2d_igemm (halo): cannot reserve %u bytes of shared memory: %s", L::TOTAL, cudaGetErrorString(e));
            return GT_ERR_CUDA;
        }
        configured = true;
    }
    int grid = gt_num_sms();
    if (grid > total_items) grid = total_items;
    conv_igemm_halo_kernel<BN, MT, SB, NBUF, MODE><<<grid, MODE == CONV_F16_EP ? NTHREADS_EP : NTHREADS, L::TOTAL, stream>>>(tmA, tmB, p, total_items);
    GT_CUDA_LAUNCH_CHECK("gt_conv2d_igemm (halo)");
    return GT_OK;
}

template <int BN, int MT, int SB, int NBUF>
int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int total_items, cudaStream_t stream) {
    switch (conv_mode(p)) {
        case CONV_F32OUT: return launch_halo_m<BN, MT, SB, NBUF, CONV_F32OUT>(tmA, tmB, p, total_items, stream);
        case CONV_F16_EP: return launch_halo_m<BN, MT, SB, NBUF, CONV_F16_EP>(tmA, tmB, p, total_items, stream);
        default: return launch_halo_m<BN, MT, SB, NBUF, CONV_F16>(tmA, tmB, p, total_items, stream);
    }
}

}  // namespace

bool gt_conv_halo2_applicable(const ConvParams& p, int maxOH, int maxOW);
int gt_launch_conv_halo2(const void* x, long long xs_n, long long xs_h, long long xs_w, int H, int W, const void* wpacked, int ntaps_total, ConvParams& p,
                         int ext_x, int ext_y, int maxOH, int maxOW, cudaStream_t stream);

int g_conv_halo_tuning = 0;   // 0: CTA-pair kernel where applicable, else as 7; 7: single-CTA kernel (Cout%256 -> BN 256 x 2 sub-tiles, one accumulator set);
                              // 2: single CTA, BN 256 x 1 sub-tile, two sets; 3: single CTA, BN 128 everywhere; 6: single CTA without the global stores (experiments)

static void pick_tile(int Cout, int Cin, int& BN, int& MT) {
    if (Cout % 256 == 0 && g_conv_halo_tuning != 3) {
        BN = 256;
        // measured (profiles/r01_conv_microbench.txt): two sub-tiles sharing the weight tile win once Cin >= 512,
        // one sub-tile with double-buffered accumulators wins below
        MT = (g_conv_halo_tuning == 2) ? 1 : (g_conv_halo_tuning == 4 ? 2 : (Cin >= 512 ? 2 : 1));
    } else if (Cout % 128 == 0) {
        BN = 128;
        MT = 2;
    } else {
        BN = 64;
        MT = 4;
    }
}

bool gt_conv_halo_applicable(const ConvParams& p, int maxOH, int maxOW) {
    if (p.in_stride != 1) return false;
    int bn, mt;
    pick_tile(p.Cout, p.Cin, bn, mt);
    // worth it only when a tile is mostly inside the image
    return maxOH >= SUB_H && maxOW >= SUB_W * mt;
}

int gt_launch_conv_halo(const void* x, long long xs_n, long long xs_h, long long xs_w, int H, int W, const void* wpacked, int ntaps_total, ConvParams& p,
                        cudaStream_t stream) {
    int BN, MT;
    pick_tile(p.Cout, p.Cin, BN, MT);
    int ext_x = 0, ext_y = 0, maxOH = 0, maxOW = 0;
    for (int i = 0; i < p.nphases; i++) {
        const ConvPhase& ph = p.ph[i];
        int dy0 = 127, dy1 = -128, dx0 = 127, dx1 = -128;
        for (int t = 0; t < ph.ntaps; t++) {
            dy0 = ph.tdy[t] < dy0 ? ph.tdy[t] : dy0;
            dy1 = ph.tdy[t] > dy1 ? ph.tdy[t] : dy1;
            dx0 = ph.tdx[t] < dx0 ? ph.tdx[t] : dx0;
            dx1 = ph.tdx[t] > dx1 ? ph.tdx[t] : dx1;
        }
        p.dy_min[i] = dy0;
        p.dx_min[i] = dx0;
        ext_y = (dy1 - dy0) > ext_y ? (dy1 - dy0) : ext_y;
        ext_x = (dx1 - dx0) > ext_x ? (dx1 - dx0) : ext_x;
        maxOH = ph.OHp > maxOH ? ph.OHp : maxOH;
        maxOW = ph.OWp > maxOW ? ph.OWp : maxOW;
    }
    p.dbg_no_store = (g_conv_halo_tuning == 6);
    GT_REQUIRE(ext_x <= 2 && ext_y <= 2, "gt_conv2d_igemm_f16 (halo): tap extent %dx%d exceeds the staged halo", ext_x, ext_y);
    // CTA-pair kernel (conv_igemm_halo2.cu): the default wherever it applies -- measured faster than the single-CTA kernel on every
    // stride-1 phase (r02e: 64ch 163 -> 141 us, 128ch 112 -> 101, 256ch 105 -> 95, 512ch 114 -> 103); tuning 7 forces the single-CTA kernel
    if (g_conv_halo_tuning != 7 && (g_conv_halo_tuning == 0 || g_conv_halo_tuning == 5) && gt_conv_halo2_applicable(p, maxOH, maxOW))
        return gt_launch_conv_halo2(x, xs_n, xs_h, xs_w, H, W, wpacked, ntaps_total, p, ext_x, ext_y, maxOH, maxOW, stream);
    p.halo_w = SUB_W * MT + ext_x;
    p.halo_h = SUB_H + ext_y;
    p.n_tiles = p.Cout / BN;
    p.tiles_w = (maxOW + SUB_W * MT - 1) / (SUB_W * MT);
    p.tiles_h = (maxOH + SUB_H - 1) / SUB_H;
    p.tiles_n = p.N;
    const long long total = (long long)p.nphases * p.N * p.tiles_h * p.tiles_w * p.n_tiles;
    GT_REQUIRE(total < (1ll << 31), "gt_conv2d_igemm_f16 (halo): too many tiles");

    gt_encode_tiled_fn encode = gt_get_encode_tiled();
    GT_REQUIRE(encode != nullptr, "gt_conv2d_igemm_f16: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t esz = 2;
    const int kel = 64;
    const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)xs_w * esz, (cuuint64_t)xs_h * esz, (cuuint64_t)xs_n * esz};
        cuuint32_t box[4] = {(cuuint32_t)kel, (cuuint32_t)p.halo_w, (cuuint32_t)p.halo_h, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, dt, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm_f16 (halo): activation tensor map rejected (CUresult %d)", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)ntaps_total};
        cuuint64_t strides[2] = {(cuuint64_t)p.Cin * esz, (cuuint64_t)p.Cin * p.Cout * esz};
        cuuint32_t box[3] = {(cuuint32_t)kel, (cuuint32_t)BN, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, dt, 3, const_cast<void*>(wpacked), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm_f16 (halo): weight tensor map rejected (CUresult %d)", (int)r);
    }
    if (BN == 256 && MT == 2) return launch_halo<256, 2, 4, 1>(tmA, tmB, p, (int)total, stream);
    if (BN == 256) return launch_halo<256, 1, 5, 2>(tmA, tmB, p, (int)total, stream);
    if (BN == 128) return launch_halo<128, 2, 6, 2>(tmA, tmB, p, (int)total, stream);
    return launch_halo<64, 4, 6, 2>(tmA, tmB, p, (int)total, stream);
}
