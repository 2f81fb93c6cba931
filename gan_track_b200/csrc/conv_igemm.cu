// conv_igemm.cu -- fp16 NHWC convolution as an implicit GEMM on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// What it replaces: the reference hands every convolution of the StyleGAN2 path to cuDNN through
// torch.nn.functional.conv2d / conv_transpose2d (OPS/conv2d_gradfix.py:37-45; OPS =
// /root/reference/src/models/stylegan3/torch_utils/ops).  The cases on the path (OPS/conv2d_resample.py:94-134 and
// their data gradients) are
//     3x3 stride 1 pad 1, 1x1            -> "plain"      one phase, taps (r - pad, s - pad)
//     3x3 stride 2 pad 0                 -> "strided"    one phase, TMA walks the input with element stride 2
//     3x3 transposed stride 2 pad 0      -> "transposed" four output-parity phases with 4 / 2 / 2 / 1 taps each
//     transposed stride 1                -> plain with taps (pad - r, pad - s)
// and every data gradient of one case is another case of the same list, so this one kernel is forward AND dgrad.
//
// GEMM view (per phase):  D[pixel, co] = sum_{tap, ci} A_tap[pixel, ci] * B_tap[co, ci]
//     M = 128 output pixels  = a box of bn x bh x bw pixels of the NHWC activation tensor
//     N = BN output channels (64 / 128 / 256)
//     K = taps x Cin, walked in k-blocks of 64 channels (128 bytes = one SWIZZLE_128B row)
// A_tap is fetched by ONE 4-D TMA box load per k-block, shifted by the tap offset; out-of-bounds rows/columns are
// zero-filled by the TMA unit, which is exactly the convolution's zero padding.  B_tap comes from weights packed as
// [tap][Cout][Cin] fp16 (pack kernel below).  Both land in shared memory K-major with the 128-byte swizzle, the
// layout tcgen05.mma reads through a shared-memory descriptor.  Accumulation is fp32 in TMEM.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warps 2-5 = epilogue
// (TMEM -> registers -> fp16 -> global, one output pixel per thread, 64 contiguous bytes per tcgen05.ld chunk).
// smem ring of STAGES {A 16 KB, B BN*128 B} guarded by full/empty mbarriers; tcgen05.commit releases a stage.
#include "conv_common.cuh"

using namespace sm100;

gt_encode_tiled_fn gt_get_encode_tiled() {
    static gt_encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (gt_encode_tiled_fn)p;
    return fn;
}

namespace {

constexpr int BM = 128;   // output pixels per tile (UMMA M)
constexpr int BK = 64;    // channels per k-block
constexpr int MAX_TAPS = CONV_MAX_TAPS;
constexpr int NTHREADS = 192;

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr uint32_t A_BYTES = BM * 128;
    static constexpr uint32_t B_BYTES = BN * 128;
    static constexpr uint32_t TILES = STAGES * (A_BYTES + B_BYTES);
    static constexpr uint32_t BARS = (2 * STAGES + 1) * 8 + 8;
    static constexpr uint32_t TOTAL = TILES + BARS + 1024;   // + slack for the manual 1024-byte alignment
};

template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(NTHREADS) conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                              const ConvParams p) {
    typedef SmemLayout<BN, STAGES> L;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * L::A_BYTES;
    uint64_t* full = (uint64_t*)(smem + L::TILES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = full + 2 * STAGES;
    uint32_t* tmem_slot = (uint32_t*)(full + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ConvPhase& ph = p.ph[blockIdx.z];

    int tile = blockIdx.x;
    const int nt = tile % p.n_tiles;
    tile /= p.n_tiles;
    const int twi = tile % p.tiles_w;
    tile /= p.tiles_w;
    const int thi = tile % p.tiles_h;
    const int tni = tile / p.tiles_h;
    const int bw_log2 = p.bw_log2, bh_log2 = p.bh_log2;
    const int ox0 = twi << bw_log2, oy0 = thi << bh_log2, n0 = tni << (7 - bw_log2 - bh_log2);
    if (ox0 >= ph.OWp || oy0 >= ph.OHp) return;   // tile outside this (smaller) phase; uniform for the CTA

    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
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

    constexpr int kel = BK;                                    // channels per 128-byte k-chunk
    const int kchunks = p.Cin / kel;
    const int nkb = ph.ntaps * kchunks;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            // channel chunks in the OUTER loop, taps inside (the order the fp16x3 route relies on: csrc/conv_f16x3.cu)
            for (int kc = 0; kc < kchunks; kc++) {
                for (int t = 0; t < ph.ntaps; t++) {
                    const int cy = oy0 * p.in_stride + ph.tdy[t], cx = ox0 * p.in_stride + ph.tdx[t];
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], L::A_BYTES + L::B_BYTES);
                    tma_load_4d(sA + stage * L::A_BYTES, &tmA, &full[stage], kc * kel, cx, cy, n0);
                    tma_load_3d(sB + stage * L::B_BYTES, &tmB, &full[stage], kc * kel, nt * BN, ph.tw[t]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(BM, BN, 0, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nkb; kb++) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + stage * L::A_BYTES), b0 = smem_u32(sB + stage * L::B_BYTES);
                // one MMA consumes 32 bytes of every row: K = 16 fp16 elements
#pragma unroll
                for (int k = 0; k < 4; k++)
                    umma_f16(tmem_base, umma_smem_desc(a0 + k * 32, 0, 1024), umma_smem_desc(b0 + k * 32, 0, 1024), idesc, (uint32_t)((kb | k) != 0));
                umma_commit(&empty[stage]);   // stage reusable once these MMAs have read it
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(tfull);   // accumulator complete
        }
        __syncwarp();
    } else {
        // Epilogue: a warp may only touch TMEM lanes [32 * (warp % 4), +32).
        const int q = warp & 3;
        const int m = q * 32 + lane;                                  // row of the tile = pixel within the box
        const int lw = m & ((1 << bw_log2) - 1);
        const int lh = (m >> bw_log2) & ((1 << bh_log2) - 1);
        const int ln = m >> (bw_log2 + bh_log2);
        const int a = oy0 + lh, b = ox0 + lw, n = n0 + ln;
        const bool valid = (a < ph.OHp) && (b < ph.OWp) && (n < p.N);
        const long long yoff = (long long)n * p.ys_n + (long long)(a * p.out_stride + ph.off_y) * p.ys_h +
                               (long long)(b * p.out_stride + ph.off_x) * p.ys_w + nt * BN + ph.y_off;
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; c++) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            if (valid) conv_store32<MODE>(p, yoff + c * 32, nt * BN + c * 32, r);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// out[t][co][ci] = w[co * s_co + ci * s_ci + r * s_r + s * s_s], t = r * KW + s
__global__ void pack_weight_kernel(const __half* __restrict__ w, long long s_co, long long s_ci, long long s_r, long long s_s, int Cout, int Cin,
                                   int KH, int KW, __half* __restrict__ out) {
    const long long total = (long long)KH * KW * Cout * Cin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        long long r_ = i / Cin;
        const int co = (int)(r_ % Cout);
        const int t = (int)(r_ / Cout);
        const int r = t / KW, s = t - r * KW;
        out[i] = w[co * s_co + ci * s_ci + r * s_r + s * s_s];
    }
}

int next_pow2_log2(int v) {
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

template <int BN, int STAGES, int MODE>
int launch_conv_m(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream) {
    typedef SmemLayout<BN, STAGES> L;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN, STAGES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
        if (e != cudaSuccess) {
            gt_set_error("gt_conv2d_igemm: cannot reserve %u bytes of shared memory: %s", L::TOTAL, cudaGetErrorString(e));
            return GT_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid((unsigned)(p.n_tiles * p.tiles_w * p.tiles_h * p.tiles_n), 1, (unsigned)p.nphases);
    conv_igemm_kernel<BN, STAGES, MODE><<<grid, NTHREADS, L::TOTAL, stream>>>(tmA, tmB, p);
    GT_CUDA_LAUNCH_CHECK("gt_conv2d_igemm");
    return GT_OK;
}

template <int BN, int STAGES>
int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream) {
    switch (conv_mode(p)) {
        case CONV_F32OUT: return launch_conv_m<BN, STAGES, CONV_F32OUT>(tmA, tmB, p, stream);
        case CONV_F16_EP: return launch_conv_m<BN, STAGES, CONV_F16_EP>(tmA, tmB, p, stream);
        default: return launch_conv_m<BN, STAGES, CONV_F16>(tmA, tmB, p, stream);
    }
}

}  // namespace

extern int g_conv_halo_tuning;
static int g_conv_variant = 0;   // 0 = auto (halo kernel where applicable), 1 = per-tap kernel only

extern "C" int gt_conv_igemm_config(int variant) {
    const int old = g_conv_variant;
    if (variant == 0 || variant == 1) {
        g_conv_variant = variant;
        g_conv_halo_tuning = 0;
    } else if (variant >= 2) {   // halo kernel with an alternative tile choice (tuning experiments)
        g_conv_variant = 0;
        g_conv_halo_tuning = variant;
    }
    return old;
}

extern "C" int gt_conv_pack_weight_f16(const void* w, long long s_co, long long s_ci, long long s_r, long long s_s, int Cout, int Cin, int KH, int KW,
                                       void* out, void* stream) {
    GT_REQUIRE(w && out, "gt_conv_pack_weight_f16: null pointer");
    GT_REQUIRE(Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "gt_conv_pack_weight_f16: bad shape");
    const long long total = (long long)KH * KW * Cout * Cin;
    const int block = 256;
    const int grid = (int)((total + block - 1) / block < 148 * 8 ? (total + block - 1) / block : 148 * 8);
    pack_weight_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __half*)w, s_co, s_ci, s_r, s_s, Cout, Cin, KH, KW, (__half*)out);
    GT_CUDA_LAUNCH_CHECK("gt_conv_pack_weight_f16");
    return GT_OK;
}

static int conv2d_igemm_impl(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y, long long ys_n,
                             long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride,
                             int pad, int transposed, int f32out, long long slab_stride, const void* ep_bias, int ep_act, float ep_alpha, float ep_gain,
                             float ep_clamp, const void* ep_add, void* stream) {
    // f32out: fp16 operands, raw fp32 accumulators stored; with slab_stride > 0 the reduction is additionally split by kernel row
    // (one phase per row of taps, phase i writing its partial sum at y + i * slab_stride elements), which bounds the number of
    // accumulator updates per stored value (see csrc/conv_f16x3.cu)
    const int esz = 2, kel = BK, al = 8;
    const int osz_al = f32out ? 4 : 8;
    GT_REQUIRE(ep_act == 0 || (!f32out && (ep_act == 1 || ep_act == 3)), "gt_conv2d_igemm: fused epilogue supports fp16 output with linear (1) or lrelu (3); got %d", ep_act);
    GT_REQUIRE(ep_bias == nullptr || (((uintptr_t)ep_bias) & 15) == 0, "gt_conv2d_igemm: epilogue bias must be 16-byte aligned");
    GT_REQUIRE(ep_add == nullptr || (ep_act != 0 && (((uintptr_t)ep_add) & 15) == 0), "gt_conv2d_igemm: the residual needs the fused epilogue and 16-byte alignment");
    GT_REQUIRE(x && wpacked && y, "gt_conv2d_igemm_f16: null pointer");
    GT_REQUIRE(N > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "gt_conv2d_igemm_f16: empty tensor");
    GT_REQUIRE(Cin % kel == 0 && Cout % 64 == 0, "gt_conv2d_igemm: Cin (%d) must be a multiple of %d and Cout (%d) of 64", Cin, kel, Cout);
    GT_REQUIRE(KH * KW <= MAX_TAPS && KH >= 1 && KW >= 1, "gt_conv2d_igemm_f16: kernel %dx%d not supported", KH, KW);
    GT_REQUIRE(stride == 1 || stride == 2, "gt_conv2d_igemm_f16: stride %d not supported", stride);
    GT_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)wpacked & 15) == 0 && ((uintptr_t)y & 15) == 0, "gt_conv2d_igemm_f16: pointers must be 16-byte aligned");
    GT_REQUIRE(xs_w % al == 0 && xs_h % al == 0 && xs_n % al == 0 && ys_w % osz_al == 0 && ys_h % osz_al == 0 && ys_n % osz_al == 0 && slab_stride % 4 == 0,
               "gt_conv2d_igemm: strides must be multiples of 16 bytes");
    GT_REQUIRE(pad >= 0 && pad < 8, "gt_conv2d_igemm_f16: pad %d not supported", pad);

    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.N = N;
    p.Cin = Cin;
    p.Cout = Cout;
    p.y = y;
    p.f32out = f32out;
    p.ep_bias = (const __half*)ep_bias;
    p.ep_add = (const __half*)ep_add;
    p.ep_act = ep_act;
    p.ep_alpha = ep_alpha;
    p.ep_gain = ep_gain;
    p.ep_clamp = ep_clamp;
    p.ys_n = ys_n;
    p.ys_h = ys_h;
    p.ys_w = ys_w;
    if (!transposed) {
        GT_REQUIRE(OH == (H + 2 * pad - KH) / stride + 1 && OW == (W + 2 * pad - KW) / stride + 1, "gt_conv2d_igemm_f16: output size mismatch");
        const bool ksplit = f32out && slab_stride > 0 && KH > 1;
        GT_REQUIRE(!ksplit || KH <= CONV_MAX_PHASES, "gt_conv2d_igemm: K-split supports at most %d kernel rows", CONV_MAX_PHASES);
        p.nphases = ksplit ? KH : 1;
        p.in_stride = stride;
        p.out_stride = 1;
        for (int r = 0; r < KH; r++) {
            ConvPhase& ph = p.ph[ksplit ? r : 0];
            ph.OHp = OH;
            ph.OWp = OW;
            ph.y_off = ksplit ? (long long)r * slab_stride : 0;
            for (int s = 0; s < KW; s++) {
                const int t = ksplit ? s : r * KW + s;
                ph.tdy[t] = (int8_t)(r - pad);
                ph.tdx[t] = (int8_t)(s - pad);
                ph.tw[t] = (int8_t)(r * KW + s);
                ph.ntaps = t + 1;
            }
        }
    } else if (stride == 1) {
        GT_REQUIRE(OH == H - 2 * pad + KH - 1 && OW == W - 2 * pad + KW - 1, "gt_conv2d_igemm_f16: output size mismatch (transposed)");
        const bool ksplit = f32out && slab_stride > 0 && KH > 1;
        GT_REQUIRE(!ksplit || KH <= CONV_MAX_PHASES, "gt_conv2d_igemm: K-split supports at most %d kernel rows", CONV_MAX_PHASES);
        p.nphases = ksplit ? KH : 1;
        p.in_stride = 1;
        p.out_stride = 1;
        for (int r = 0; r < KH; r++) {
            ConvPhase& ph = p.ph[ksplit ? r : 0];
            ph.OHp = OH;
            ph.OWp = OW;
            ph.y_off = ksplit ? (long long)r * slab_stride : 0;
            for (int s = 0; s < KW; s++) {
                const int t = ksplit ? s : r * KW + s;
                ph.tdy[t] = (int8_t)(pad - r);
                ph.tdx[t] = (int8_t)(pad - s);
                ph.tw[t] = (int8_t)(r * KW + s);
                ph.ntaps = t + 1;
            }
        }
    } else {
        GT_REQUIRE(pad == 0, "gt_conv2d_igemm_f16: transposed stride-2 convolution supports pad 0 only");
        GT_REQUIRE(OH == (H - 1) * 2 + KH && OW == (W - 1) * 2 + KW, "gt_conv2d_igemm_f16: output size mismatch (transposed stride 2)");
        p.nphases = 4;
        p.in_stride = 1;
        p.out_stride = 2;
        for (int py = 0; py < 2; py++)
            for (int px = 0; px < 2; px++) {
                ConvPhase& ph = p.ph[py * 2 + px];
                ph.OHp = (OH - py + 1) / 2;
                ph.OWp = (OW - px + 1) / 2;
                ph.off_y = py;
                ph.off_x = px;
                int t = 0;
                for (int r = py; r < KH; r += 2)
                    for (int s = px; s < KW; s += 2) {
                        ph.tdy[t] = (int8_t)(-(r - py) / 2);
                        ph.tdx[t] = (int8_t)(-(s - px) / 2);
                        ph.tw[t] = (int8_t)(r * KW + s);
                        t++;
                    }
                ph.ntaps = t;
            }
    }
    int maxOH = 0, maxOW = 0;
    for (int i = 0; i < p.nphases; i++) {
        GT_REQUIRE(p.ph[i].ntaps >= 1, "gt_conv2d_igemm_f16: empty phase");
        maxOH = p.ph[i].OHp > maxOH ? p.ph[i].OHp : maxOH;
        maxOW = p.ph[i].OWp > maxOW ? p.ph[i].OWp : maxOW;
    }
    // Tile box: up to 16 pixels wide, then as tall as fits, the rest of the 128 rows across images.
    int bw_log2 = next_pow2_log2(maxOW);
    if (bw_log2 > 4) bw_log2 = 4;
    int bh_log2 = next_pow2_log2(maxOH);
    if (bh_log2 > 7 - bw_log2) bh_log2 = 7 - bw_log2;
    const int bn_log2 = 7 - bw_log2 - bh_log2;
    const int bw = 1 << bw_log2, bh = 1 << bh_log2, bn = 1 << bn_log2;
    p.bw_log2 = bw_log2;
    p.bh_log2 = bh_log2;
    p.tiles_w = (maxOW + bw - 1) / bw;
    p.tiles_h = (maxOH + bh - 1) / bh;
    p.tiles_n = (N + bn - 1) / bn;

    if (g_conv_variant != 1 && gt_conv_rows_applicable(p, H, W))
        return gt_launch_conv_rows(x, xs_n, xs_h, xs_w, H, W, wpacked, KH * KW, p, (cudaStream_t)stream);
    if (g_conv_variant != 1 && p.in_stride == 1 && gt_conv_halo_applicable(p, maxOH, maxOW))
        return gt_launch_conv_halo(x, xs_n, xs_h, xs_w, H, W, wpacked, KH * KW, p, (cudaStream_t)stream);

    gt_encode_tiled_fn encode = gt_get_encode_tiled();
    GT_REQUIRE(encode != nullptr, "gt_conv2d_igemm_f16: cuTensorMapEncodeTiled is not available from this driver");
    const int BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
    p.n_tiles = Cout / BN;

    const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)xs_w * esz, (cuuint64_t)xs_h * esz, (cuuint64_t)xs_n * esz};
        cuuint32_t box[4] = {(cuuint32_t)kel, (cuuint32_t)(bw * p.in_stride), (cuuint32_t)(bh * p.in_stride), (cuuint32_t)bn};
        cuuint32_t estr[4] = {1, (cuuint32_t)p.in_stride, (cuuint32_t)p.in_stride, 1};
        CUresult r = encode(&tmA, dt, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm_f16: activation tensor map rejected (CUresult %d)", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)(KH * KW)};
        cuuint64_t strides[2] = {(cuuint64_t)Cin * esz, (cuuint64_t)Cin * Cout * esz};
        cuuint32_t box[3] = {(cuuint32_t)kel, (cuuint32_t)BN, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, dt, 3, const_cast<void*>(wpacked), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GT_REQUIRE(r == CUDA_SUCCESS, "gt_conv2d_igemm_f16: weight tensor map rejected (CUresult %d)", (int)r);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (BN == 256) return launch_conv<256, 4>(tmA, tmB, p, st);
    if (BN == 128) return launch_conv<128, 3>(tmA, tmB, p, st);
    return launch_conv<64, 4>(tmA, tmB, p, st);
}

extern "C" int gt_conv2d_igemm_f16(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y, long long ys_n,
                                   long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride,
                                   int pad, int transposed, void* stream) {
    return conv2d_igemm_impl(x, xs_n, xs_h, xs_w, wpacked, y, ys_n, ys_h, ys_w, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, transposed, 0, 0, nullptr, 0,
                             0.f, 1.f, -1.f, nullptr, stream);
}

// fp16 operands, fp32 output (raw accumulators): the tensor-core half of the fp16x3 route for the fp32 layers (csrc/conv_f16x3.cu).
// slab_stride > 0 and a kernel with several rows: one partial sum per kernel row at y + row * slab_stride (elements); the caller adds them.
extern "C" int gt_conv2d_igemm_f16_f32out(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y, long long ys_n,
                                          long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride,
                                          int pad, int transposed, long long slab_stride, void* stream) {
    return conv2d_igemm_impl(x, xs_n, xs_h, xs_w, wpacked, y, ys_n, ys_h, ys_w, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, transposed, 1, slab_stride,
                             nullptr, 0, 0.f, 1.f, -1.f, nullptr, stream);
}

// the same convolution with the layer's bias_act fused into the epilogue: y = clamp(act(conv + bias) * gain)
extern "C" int gt_conv2d_igemm_f16_bias_act(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y, long long ys_n,
                                            long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW,
                                            int stride, int pad, int transposed, const void* bias, int act, float alpha, float gain, float clamp,
                                            void* stream) {
    GT_REQUIRE(act == 1 || act == 3, "gt_conv2d_igemm_f16_bias_act: act must be linear (1) or lrelu (3); got %d", act);
    return conv2d_igemm_impl(x, xs_n, xs_h, xs_w, wpacked, y, ys_n, ys_h, ys_w, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, transposed, 0, 0, bias, act,
                             alpha, gain, clamp, nullptr, stream);
}

// ... plus a residual: y = fp16(clamp(act(conv + bias) * gain)) + addend, addend fp16 with exactly the shape and strides of y
// (DiscriminatorBlock.forward's `y.add_(x)`, S3/training/networks_stylegan2.py:636, folded into the skip convolution)
extern "C" int gt_conv2d_igemm_f16_bias_act_add(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y, long long ys_n,
                                                long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW,
                                                int stride, int pad, int transposed, const void* bias, int act, float alpha, float gain, float clamp,
                                                const void* addend, void* stream) {
    GT_REQUIRE(act == 1 || act == 3, "gt_conv2d_igemm_f16_bias_act_add: act must be linear (1) or lrelu (3); got %d", act);
    GT_REQUIRE(addend != nullptr, "gt_conv2d_igemm_f16_bias_act_add: null addend");
    return conv2d_igemm_impl(x, xs_n, xs_h, xs_w, wpacked, y, ys_n, ys_h, ys_w, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, transposed, 0, 0, bias, act,
                             alpha, gain, clamp, addend, stream);
}
