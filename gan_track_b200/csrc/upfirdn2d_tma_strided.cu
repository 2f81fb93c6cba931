// upfirdn2d_tma_strided.cu -- the 4x4 FIR with a factor-2 resampling (up = 2 or down = 2) on channels-last tensors,
// staged through shared memory by TMA.
//
// These are the upfirdn2d calls of the discriminator's skip branch (down = 2 before the 1x1 convolution,
// OPS/conv2d_resample.py:94-97; OPS = /root/reference/src/models/stylegan3/torch_utils/ops) and of its backward
// (up = 2, OPS/upfirdn2d.py:256-266); the reference runs them on the kernels of OPS/upfirdn2d.cu:97-200
// (small<T,1,1,2,2,4,4,...> / small<T,2,2,1,1,4,4,...>).  Same structure as upfirdn2d_tma.cu: persistent CTAs, a ring of
// TMA boxes of one 128-byte channel chunk whose out-of-bounds pixels are zero-filled (= the op's zero padding), 16-byte
// vectors, packed fp32x2 arithmetic; taps that fall on inserted zeros are pruned at compile time.
//
//   down = 2: 8x8 output pixels from an 18x18 input box; a thread owns one output column x 2 output rows x one vector.
//   up   = 2: 16x16 output pixels from a 10x10 input box; a thread owns 2x2 output pixels (the four phases) x one
//             vector, twice per tile.  Output (oy, ox) uses taps k with (o + k - pad0) even; relative to the box origin
//             floor((o0 - pad0) / 2) the sample index is q + j with k = 2 j - (pad0 & 1) - c, c = o & 1, j in 0..2.
#include "gt_common.cuh"
#include "gt_sm100.cuh"
#include "hot_act.cuh"

using namespace sm100;

namespace {

constexpr int FW = 4, FH = 4;

struct FirSParams {
    void* y;
    long long ys_n, ys_h, ys_w;
    int N, OH, OW, chunks, tiles_x, tiles_y;
    int padx0, pady0;
    const float* f;        // device taps [FH,FW] with element strides fs_h / fs_w
    long long fs_h, fs_w;
    int flip;
    float gain;
};

__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

struct TileIdx {
    int ch, tx, ty, n;
};
__device__ __forceinline__ TileIdx decode_tile(int tile, const FirSParams& p) {
    TileIdx r;
    int t = tile;
    r.ch = t % p.chunks;
    t /= p.chunks;
    r.tx = t % p.tiles_x;
    t /= p.tiles_x;
    r.ty = t % p.tiles_y;
    r.n = t / p.tiles_y;
    return r;
}

// Common prologue: correlation taps with the gain folded in (g = f if flip else f reversed, OPS/upfirdn2d.py:196-199).
__device__ __forceinline__ void load_taps(float* sf, const FirSParams& p) {
    if (threadIdx.x < FH * FW) {
        const int ky = threadIdx.x / FW, kx = threadIdx.x - ky * FW;
        const int sy = p.flip ? ky : FH - 1 - ky, sx = p.flip ? kx : FW - 1 - kx;
        sf[threadIdx.x] = p.f[sy * p.fs_h + sx * p.fs_w] * p.gain;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// down = 2
// ---------------------------------------------------------------------------------------------------------------------
namespace d2 {
constexpr int TW = 8, TH = 8, STAGES = 2;
constexpr int BOX_W = 2 * (TW - 1) + FW, BOX_H = 2 * (TH - 1) + FH;   // 18 x 18
constexpr int STAGE_BYTES = BOX_W * BOX_H * 128;
constexpr int ROWS = 2;                                               // output rows per thread
constexpr int SMEM = STAGES * STAGE_BYTES + STAGES * 8 + 128;
}  // namespace d2

template <class T>
__global__ void __launch_bounds__(256, 2) upfirdn2d_tma_down2_kernel(const __grid_constant__ CUtensorMap tmX, const FirSParams p, const int total_tiles) {
    using namespace d2;
    typedef hot::Lanes<T> L;
    constexpr int NP = L::NP;
    constexpr int VEC = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    __shared__ float sf[FH * FW];
    load_taps(sf, p);
    const int tid = threadIdx.x;
    const int cv = tid & 7, cp = (tid >> 3) & 7, rg = tid >> 6;
    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int tile, int stage) {
        const TileIdx t = decode_tile(tile, p);
        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        tma_load_4d(smem + stage * STAGE_BYTES, &tmX, &full[stage], t.ch * (128 / (int)sizeof(T)), t.tx * TW * 2 - p.padx0, t.ty * TH * 2 - p.pady0, t.n);
    };
    const int first = blockIdx.x, step = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            const int tile = first + s * step;
            if (tile < total_tiles) issue(tile, s);
        }
    }
    float w[FH][FW];
#pragma unroll
    for (int a = 0; a < FH; a++)
#pragma unroll
        for (int b = 0; b < FW; b++) w[a][b] = sf[a * FW + b];

    int it = 0;
    for (int tile = first; tile < total_tiles; tile += step, it++) {
        const int stage = it % STAGES;
        mbar_wait(&full[stage], (uint32_t)((it / STAGES) & 1));
        const TileIdx t = decode_tile(tile, p);
        const int ox = t.tx * TW + cp;
        const int oy0 = t.ty * TH + rg * ROWS;
        const uint32_t sp = smem_u32(smem) + stage * STAGE_BYTES + ((rg * ROWS * 2) * BOX_W + cp * 2) * 128 + cv * 16;
        T* yp = (T*)p.y + (long long)t.n * p.ys_n + (long long)ox * p.ys_w + (long long)t.ch * (128 / (int)sizeof(T)) + cv * VEC;

        float2 acc[ROWS][NP];
#pragma unroll
        for (int o = 0; o < ROWS; o++)
#pragma unroll
            for (int k = 0; k < NP; k++) acc[o][k] = splat2(0.f);
        // staged rows tr = 0 .. 2 ROWS + FH - 3 of this thread's strip; output row o takes tap ky = tr - 2 o
#pragma unroll
        for (int tr = 0; tr < 2 * (ROWS - 1) + FH; tr++) {
            float2 v[FW][NP];
#pragma unroll
            for (int j = 0; j < FW; j++) {
                Vec16<T> raw;
                *reinterpret_cast<uint4*>(raw.v) = lds128(sp + (tr * BOX_W + j) * 128);
#pragma unroll
                for (int k = 0; k < NP; k++) v[j][k] = L::get(raw, k);
            }
#pragma unroll
            for (int o = 0; o < ROWS; o++) {
                const int ky = tr - 2 * o;
                if (ky >= 0 && ky < FH) {
#pragma unroll
                    for (int kx = 0; kx < FW; kx++)
#pragma unroll
                        for (int k = 0; k < NP; k++) acc[o][k] = __ffma2_rn(v[kx][k], splat2(w[ky][kx]), acc[o][k]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < ROWS; o++) {
            Vec16<T> out;
#pragma unroll
            for (int k = 0; k < NP; k++) L::set(out, k, acc[o][k]);
            if (ox < p.OW && oy0 + o < p.OH) st16_stream(yp + (long long)(oy0 + o) * p.ys_h, out);
        }
        __syncthreads();                       // every thread is done reading this stage
        if (tid == 0) {
            const int next = tile + STAGES * step;
            if (next < total_tiles) issue(next, stage);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// up = 2
// ---------------------------------------------------------------------------------------------------------------------
namespace u2 {
constexpr int TW = 16, TH = 16, STAGES = 4;
constexpr int BOX_W = TW / 2 + 2, BOX_H = TH / 2 + 2;                 // 10 x 10
constexpr int STAGE_BYTES = BOX_W * BOX_H * 128;
constexpr int SMEM = STAGES * STAGE_BYTES + STAGES * 8 + 128;
}  // namespace u2

// PX / PY = parity of padx0 / pady0 (decides which taps meet which output phase).
template <class T, int PX, int PY>
__global__ void __launch_bounds__(256, 2) upfirdn2d_tma_up2_kernel(const __grid_constant__ CUtensorMap tmX, const FirSParams p, const int total_tiles) {
    using namespace u2;
    typedef hot::Lanes<T> L;
    constexpr int NP = L::NP;
    constexpr int VEC = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    __shared__ float sf[FH * FW];
    load_taps(sf, p);
    const int tid = threadIdx.x;
    const int cv = tid & 7, qx = (tid >> 3) & 7, qyb = tid >> 6;
    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int bx = (-p.padx0) >> 1, by = (-p.pady0) >> 1;      // floor(-pad0 / 2): box origin relative to the tile's first sample
    auto issue = [&](int tile, int stage) {
        const TileIdx t = decode_tile(tile, p);
        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        tma_load_4d(smem + stage * STAGE_BYTES, &tmX, &full[stage], t.ch * (128 / (int)sizeof(T)), t.tx * (TW / 2) + bx, t.ty * (TH / 2) + by, t.n);
    };
    const int first = blockIdx.x, step = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            const int tile = first + s * step;
            if (tile < total_tiles) issue(tile, s);
        }
    }
    float w[FH][FW];
#pragma unroll
    for (int a = 0; a < FH; a++)
#pragma unroll
        for (int b = 0; b < FW; b++) w[a][b] = sf[a * FW + b];

    int it = 0;
    for (int tile = first; tile < total_tiles; tile += step, it++) {
        const int stage = it % STAGES;
        mbar_wait(&full[stage], (uint32_t)((it / STAGES) & 1));
        const TileIdx t = decode_tile(tile, p);
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const int qy = qyb + 4 * pass;
            const int ox = t.tx * TW + 2 * qx, oy = t.ty * TH + 2 * qy;
            const uint32_t sp = smem_u32(smem) + stage * STAGE_BYTES + (qy * BOX_W + qx) * 128 + cv * 16;
            T* yp = (T*)p.y + (long long)t.n * p.ys_n + (long long)oy * p.ys_h + (long long)ox * p.ys_w + (long long)t.ch * (128 / (int)sizeof(T)) + cv * VEC;
            float2 acc[2][2][NP];
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++)
#pragma unroll
                    for (int k = 0; k < NP; k++) acc[r][c][k] = splat2(0.f);
#pragma unroll
            for (int jy = 0; jy < 3; jy++) {
                float2 v[3][NP];
#pragma unroll
                for (int jx = 0; jx < 3; jx++) {
                    Vec16<T> raw;
                    *reinterpret_cast<uint4*>(raw.v) = lds128(sp + (jy * BOX_W + jx) * 128);
#pragma unroll
                    for (int k = 0; k < NP; k++) v[jx][k] = L::get(raw, k);
                }
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int ky = 2 * jy - PY - r;
                    if (ky >= 0 && ky < FH) {
#pragma unroll
                        for (int c = 0; c < 2; c++)
#pragma unroll
                            for (int jx = 0; jx < 3; jx++) {
                                const int kx = 2 * jx - PX - c;
                                if (kx >= 0 && kx < FW) {
#pragma unroll
                                    for (int k = 0; k < NP; k++) acc[r][c][k] = __ffma2_rn(v[jx][k], splat2(w[ky][kx]), acc[r][c][k]);
                                }
                            }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    Vec16<T> out;
#pragma unroll
                    for (int k = 0; k < NP; k++) L::set(out, k, acc[r][c][k]);
                    if (ox + c < p.OW && oy + r < p.OH) st16_stream(yp + (long long)r * p.ys_h + (long long)c * p.ys_w, out);
                }
        }
        __syncthreads();                       // every thread is done reading this stage
        if (tid == 0) {
            const int next = tile + STAGES * step;
            if (next < total_tiles) issue(next, stage);
        }
    }
}

template <class K>
int launch_strided(K kernel, int smem, bool& configured, const CUtensorMap& tm, const FirSParams& p, int total_tiles, cudaStream_t st) {
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            gt_set_error("gt_upfirdn2d(tma strided): cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return GT_ERR_CUDA;
        }
        configured = true;
    }
    int grid = gt_num_sms() * 2;
    if (grid > total_tiles) grid = total_tiles;
    kernel<<<grid, 256, smem, st>>>(tm, p, total_tiles);
    GT_CUDA_LAUNCH_CHECK("gt_upfirdn2d(tma strided)");
    return GT_OK;
}

template <class T>
int launch_up2(int px, int py, const CUtensorMap& tm, const FirSParams& p, int total, cudaStream_t st) {
    static bool cfg[4] = {false, false, false, false};
    switch (px * 2 + py) {
        case 0: return launch_strided(upfirdn2d_tma_up2_kernel<T, 0, 0>, u2::SMEM, cfg[0], tm, p, total, st);
        case 1: return launch_strided(upfirdn2d_tma_up2_kernel<T, 0, 1>, u2::SMEM, cfg[1], tm, p, total, st);
        case 2: return launch_strided(upfirdn2d_tma_up2_kernel<T, 1, 0>, u2::SMEM, cfg[2], tm, p, total, st);
        default: return launch_strided(upfirdn2d_tma_up2_kernel<T, 1, 1>, u2::SMEM, cfg[3], tm, p, total, st);
    }
}

template <class T>
int launch_down2(const CUtensorMap& tm, const FirSParams& p, int total, cudaStream_t st) {
    static bool cfg = false;
    return launch_strided(upfirdn2d_tma_down2_kernel<T>, d2::SMEM, cfg, tm, p, total, st);
}

}  // namespace

// Returns GT_OK when a TMA kernel took the call, -1 when the call is outside the coverage (the caller then uses the
// generic kernels of upfirdn2d.cu), another code on error.  Coverage: 4x4 taps, (up, down) = (2, 1) or (1, 2) in both
// axes, fp16 / fp32, channel stride 1, C a multiple of one 128-byte chunk, 16-byte aligned pointers and strides.
int gt_upfirdn2d_try_tma_strided(const void* x, const float* f, long long fs_h, long long fs_w, int flip, float gain, void* y, int dtype, int N, int C, int H, int W, long long xs_n,
                                 long long xs_h, long long xs_w, int OH, int OW, long long ys_n, long long ys_h, long long ys_w, int up, int down, int padx0, int pady0, cudaStream_t st) {
    const int esz = dtype == GT_F16 ? 2 : 4;
    const int chunk = 128 / esz;
    if ((dtype != GT_F16 && dtype != GT_F32) || C % chunk != 0) return -1;
    if (!((up == 2 && down == 1) || (up == 1 && down == 2))) return -1;
    if ((((uintptr_t)x) & 15) || (((uintptr_t)y) & 15)) return -1;
    if ((xs_w * esz) % 16 || (xs_h * esz) % 16 || (xs_n * esz) % 16 || (ys_w * esz) % 16 || (ys_h * esz) % 16 || (ys_n * esz) % 16) return -1;
    gt_encode_tiled_fn encode = gt_get_encode_tiled();
    if (!encode) return -1;
    const int tw = up == 2 ? u2::TW : d2::TW, th = up == 2 ? u2::TH : d2::TH;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)xs_w * esz, (cuuint64_t)xs_h * esz, (cuuint64_t)xs_n * esz};
    cuuint32_t box[4] = {(cuuint32_t)chunk, (cuuint32_t)(up == 2 ? u2::BOX_W : d2::BOX_W), (cuuint32_t)(up == 2 ? u2::BOX_H : d2::BOX_H), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, dtype == GT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(x), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -1;
    FirSParams p;
    memset(&p, 0, sizeof(p));
    p.y = y;
    p.ys_n = ys_n;
    p.ys_h = ys_h;
    p.ys_w = ys_w;
    p.N = N;
    p.OH = OH;
    p.OW = OW;
    p.chunks = C / chunk;
    p.tiles_x = (OW + tw - 1) / tw;
    p.tiles_y = (OH + th - 1) / th;
    p.padx0 = padx0;
    p.pady0 = pady0;
    p.f = f;
    p.fs_h = fs_h;
    p.fs_w = fs_w;
    p.flip = flip;
    p.gain = gain;
    const long long total = (long long)N * p.tiles_y * p.tiles_x * p.chunks;
    if (total <= 0 || total >= (1ll << 31)) return -1;
    if (up == 2) return dtype == GT_F16 ? launch_up2<__half>(padx0 & 1, pady0 & 1, tm, p, (int)total, st) : launch_up2<float>(padx0 & 1, pady0 & 1, tm, p, (int)total, st);
    return dtype == GT_F16 ? launch_down2<__half>(tm, p, (int)total, st) : launch_down2<float>(tm, p, (int)total, st);
}
