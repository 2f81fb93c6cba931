// conv_igemm_halo2.cu -- CTA-pair (tcgen05 cta_group::2) variant of the halo-staged implicit-GEMM convolution.
//
// Why: with both operands in shared memory the measured MMA pacing of the single-CTA kernel follows the operand fetch,
// T ~ (M + N) * K * 2 B / 64 B/clk: N = 64 layers run at a third of the tensor peak, N = 256 at two thirds.
// A CTA pair issues ONE M = 256 MMA over the two CTAs' accumulators; each CTA stages its own 128-pixel sub-tile (A) but only
// HALF of the weight tile (B), so the per-SM operand traffic drops from (128 + N) to (128 + N/2) rows per K step -- and the
// weight TMA traffic per SM halves as well.
//
// Structure (same item walk, halo staging and warp roles as conv_igemm_halo.cu):
//   * cluster of 2 CTAs = one 16-pixel-wide x 16-tall output tile; CTA rank r owns the 8-pixel-wide sub-tile r.
//   * both CTAs run a TMA producer lane (own A halo box, own half of B) whose loads complete on the LEADER's `full` mbarriers
//     (cp.async.bulk.tensor ... cta_group::2 with the peer bit of the barrier address cleared);
//   * the leader's MMA lane issues tcgen05.mma.cta_group::2 and releases stages / publishes accumulators with
//     tcgen05.commit ... multicast::cluster to BOTH CTAs' `empty` / `t_full` mbarriers;
//   * both CTAs' epilogue warps drain their own TMEM half and arrive on the leader's `t_empty` (count 8).
//   TMEM: 512 columns allocated pair-wide, two accumulator sets of BN columns.
#include "conv_common.cuh"

using namespace sm100;

namespace {

constexpr int NTHREADS = 192;
constexpr int NTHREADS_EP = 320;     // fused bias_act: eight epilogue warps (two per TMEM lane quarter, alternating 32-column chunks); with four the
                                     // extra per-element work makes the narrow layers epilogue-paced (128 channels: 158 us separate, 167 us fused)
constexpr int SUB_W = 8, SUB_H = 16;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // shared::cluster address -> same offset in the pair's even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {       // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

template <int BN, int MT, int SB>
struct Halo2Smem {
    static constexpr uint32_t A_BYTES = (((SUB_W * MT + 2) * (SUB_H + 2) * 128 + 1023) / 1024) * 1024;
    static constexpr uint32_t B_BYTES = (BN / 2) * 128;          // this CTA's half of the weight tile
    static constexpr uint32_t TILES = 2 * A_BYTES + SB * B_BYTES;
    static constexpr uint32_t NBARS = 4 + 2 * SB + 4;
    static constexpr uint32_t TOTAL = TILES + NBARS * 8 + 16 + 1024;
};

struct Item2 {
    int phase, nt, ox0, oy0, n;
    bool valid;
};
__device__ __forceinline__ Item2 decode_item2(const ConvParams& p, int item, int pair_w) {
    Item2 it;
    it.nt = item % p.n_tiles;
    int t = item / p.n_tiles;
    const int twi = t % p.tiles_w;
    t /= p.tiles_w;
    const int thi = t % p.tiles_h;
    t /= p.tiles_h;
    it.n = t % p.N;
    it.phase = t / p.N;
    it.ox0 = twi * pair_w;
    it.oy0 = thi * SUB_H;
    it.valid = it.ox0 < p.ph[it.phase].OWp && it.oy0 < p.ph[it.phase].OHp;
    return it;
}

template <int BN, int MT, int SB, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MODE == CONV_F16_EP ? NTHREADS_EP : NTHREADS, 1)
    conv_igemm_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p, const int total_items) {
    typedef Halo2Smem<BN, MT, SB> L;
    constexpr int CTA_W = SUB_W * MT, PAIR_W = 2 * CTA_W;     // pixels per CTA / per pair along x
    constexpr uint32_t ACC_COLS = MT * BN;
    static_assert(2 * ACC_COLS <= 512, "two accumulator sets must fit the tensor memory");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + 2 * L::A_BYTES;
    uint64_t* a_full = (uint64_t*)(smem + L::TILES);
    uint64_t* a_empty = a_full + 2;
    uint64_t* b_full = a_full + 4;
    uint64_t* b_empty = b_full + SB;
    uint64_t* t_full = b_empty + SB;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], 1);       // leader's producer arrive + both CTAs' bytes
            mbar_init(&a_empty[i], 1);      // one multicast commit
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], MODE == CONV_F16_EP ? 16 : 8);      // the epilogue warps of both CTAs (only the leader's is waited on)
        }
        for (int i = 0; i < SB; i++) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc2(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // the peer's barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int kchunks = p.Cin / 64;
    const uint32_t a_bytes = (uint32_t)(p.halo_w * p.halo_h) * 128u;
    const uint32_t row_pitch = (uint32_t)p.halo_w * 128u;

    if (warp == 0) {
        if (elect_one()) {
            uint32_t a_it = 0, b_it = 0;
            for (int item = cluster_id; item < total_items; item += num_clusters) {
                const Item2 it = decode_item2(p, item, PAIR_W);
                if (!it.valid) continue;
                const ConvPhase& ph = p.ph[it.phase];
                const int cx = it.ox0 + (int)rank * CTA_W + p.dx_min[it.phase], cy = it.oy0 + p.dy_min[it.phase];
                for (int kc = 0; kc < kchunks; kc++) {
                    const uint32_t ab = a_it & 1;
                    mbar_wait(&a_empty[ab], ((a_it >> 1) & 1) ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(&a_full[ab], 2 * a_bytes);
                    tma_load_4d_2sm(sA + ab * L::A_BYTES, &tmA, &a_full[ab], kc * 64, cx, cy, it.n);
                    a_it++;
                    for (int t = 0; t < ph.ntaps; t++) {
                        const uint32_t bs = b_it % SB;
                        mbar_wait(&b_empty[bs], ((b_it / SB) & 1) ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(&b_full[bs], 2 * L::B_BYTES);
                        tma_load_3d_2sm(sB + bs * L::B_BYTES, &tmB, &b_full[bs], kc * 64, it.nt * BN + (int)rank * (BN / 2), ph.tw[t]);
                        b_it++;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc(256, BN, 0, 0, 0);
            uint32_t a_it = 0, b_it = 0, t_it = 0;
            for (int item = cluster_id; item < total_items; item += num_clusters) {
                const Item2 it = decode_item2(p, item, PAIR_W);
                if (!it.valid) continue;
                const ConvPhase& ph = p.ph[it.phase];
                const int dy0 = p.dy_min[it.phase], dx0 = p.dx_min[it.phase];
                const uint32_t buf = t_it & 1;
                mbar_wait(&t_empty[buf], ((t_it >> 1) & 1) ^ 1);
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
                            for (int k = 0; k < 4; k++)
                                umma_f16_2sm(acc + j * BN, umma_smem_desc(at + j * (SUB_W * 128) + k * 32, 0, row_pitch), umma_smem_desc(b0 + k * 32, 0, 1024),
                                             idesc, (uint32_t)((kc | t | k) != 0));
                        }
                        umma_commit_2sm(&b_empty[bs]);
                        b_it++;
                    }
                    umma_commit_2sm(&a_empty[ab]);
                    a_it++;
                }
                umma_commit_2sm(&t_full[buf]);
                t_it++;
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int lw = m & (SUB_W - 1), lh = m >> 3;
        constexpr int CSETS = MODE == CONV_F16_EP ? 2 : 1;            // warp sets sharing a lane quarter take alternating column chunks
        const int cset = MODE == CONV_F16_EP ? ((warp - 2) >> 2) : 0;
        uint32_t t_it = 0;
        for (int item = cluster_id; item < total_items; item += num_clusters) {
            const Item2 it = decode_item2(p, item, PAIR_W);
            if (!it.valid) continue;
            const ConvPhase& ph = p.ph[it.phase];
            const uint32_t buf = t_it & 1;
            mbar_wait(&t_full[buf], (t_it >> 1) & 1);
            tc_fence_after();
            const int a = it.oy0 + lh;
#pragma unroll 1
            for (int j = 0; j < MT; j++) {
                const int b = it.ox0 + (int)rank * CTA_W + j * SUB_W + lw;
                const bool valid = (a < ph.OHp) && (b < ph.OWp);
                const long long yoff = (long long)it.n * p.ys_n + (long long)(a * p.out_stride + ph.off_y) * p.ys_h +
                                       (long long)(b * p.out_stride + ph.off_x) * p.ys_w + it.nt * BN;
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
            if (lane == 0) mbar_arrive_leader(&t_empty[buf]);
            t_it++;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // nobody leaves (or frees TMEM) while the pair still uses its shared memory / TMEM
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
    }
}

template <int BN, int MT, int SB, int MODE>
int launch_halo2_m(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int total_items, cudaStream_t stream) {
    typedef Halo2Smem<BN, MT, SB> L;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_halo2_kernel<BN, MT, SB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
        if (e != cudaSuccess) {
            gt_set_error("gt_conv2d_igemm (halo, CTA pair): cannot reserve %u bytes of shared memory: %s", L::TOTAL, cudaGetErrorString(e));
            return GT_ERR_CUDA;
        }
        configured = true;
    }
    int clusters = gt_num_sms() / 2;
    if (clusters > total_items) clusters = total_items;
    conv_igemm_halo2_kernel<BN, MT, SB, MODE><<<2 * clusters, MODE == CONV_F16_EP ? NTHREADS_EP : NTHREADS, L::TOTAL, stream>>>(tmA, tmB, p, total_items);
    GT_CUDA_LAUNCH_CHECK("gt_conv2d_igemm (halo, CTA pair)");
    return GT_OK;
}

template <int BN, int MT, int SB>
int launch_halo2(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int total_items, cudaStream_t stream) {
    if (conv_mode(p) == CONV_F16_EP) return launch_halo2_m<BN, MT, SB, CONV_F16_EP>(tmA, tmB, p, total_items, stream);
    return launch_halo2_m<BN, MT, SB, CONV_F16>(tmA, tmB, p, total_items, stream);
}

}  // namespace

// fp16 output modes (plain and fused bias_act epilogue); the caller (gt_launch_conv_halo) has filled dy_min / dx_min and checked the tap extent.
static int pair_mt(int Cout) { return Cout % 256 == 0 ? 1 : (Cout % 128 == 0 ? 2 : 4); }

bool gt_conv_halo2_applicable(const ConvParams& p, int maxOH, int maxOW) {
    if (p.nphases == 1 && p.ph[0].ntaps == 1) return false;      // 1x1: nothing to share between taps, the single-CTA kernel is faster (34 vs 38 us)
    return p.in_stride == 1 && conv_mode(p) != CONV_F32OUT && maxOH >= SUB_H && maxOW >= 2 * SUB_W * pair_mt(p.Cout);
}

int gt_launch_conv_halo2(const void* x, long long xs_n, long long xs_h, long long xs_w, int H, int W, const void* wpacked, int ntaps_total, ConvParams& p,
                         int ext_x, int ext_y, int maxOH, int maxOW, cudaStream_t stream) {
    const int BN = (p.Cout % 256 == 0) ? 256 : (p.Cout % 128 == 0 ? 128 : 64);
    const int MT = pair_mt(p.Cout);
    p.halo_w = SUB_W * MT + ext_x;
    p.halo_h = SUB_H + ext_y;
    p.n_tiles = p.Cout / BN;
    p.tiles_w = (maxOW + 2 * SUB_W * MT - 1) / (2 * SUB_W * MT);
    p.tiles_h = (maxOH + SUB_H - 1) / SUB_H;
    p.tiles_n = p.N;
    const long long total = (long long)p.nphases * p.N * p.tiles_h * p.tiles_w * p.n_tiles;
    GT_REQUIRE(total < (1ll << 31), "gt_conv2d_igemm (halo, CTA pair): too many tiles");
    gt_encode_tiled_fn encode = gt_get_encode_tiled();
    GT_REQUIRE(encode != nullptr, "gt_conv2d_igemm: cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)xs_w * 2, (cuuint64_t)xs_h * 2, (cuuint64_t)xs_n * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.halo_w, (cuuint32_t)p.halo_h, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm (halo, CTA pair): activation tensor map rejected (CUresult %d)", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)ntaps_total};
        cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Cin * p.Cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)(BN / 2), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(wpacked), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm (halo, CTA pair): weight tensor map rejected (CUresult %d)", (int)r);
    }
    if (BN == 256) return launch_halo2<256, 1, 10>(tmA, tmB, p, (int)total, stream);
    if (BN == 128) return launch_halo2<128, 2, 12>(tmA, tmB, p, (int)total, stream);
    return launch_halo2<64, 4, 10>(tmA, tmB, p, (int)total, stream);
}
