// upfirdn2d_tma.cu -- the 4x4 FIR blur (up = down = 1) on channels-last tensors, staged through shared memory by TMA.
//
// This is the upfirdn2d call that carries the traffic on the StyleGAN2 path: the blur after every transposed
// up-convolution and before every strided down-convolution (OPS/conv2d_resample.py:106-109, 112-126; OPS =
// /root/reference/src/models/stylegan3/torch_utils/ops; reference kernel OPS/upfirdn2d.cu:97-200, variant
// small<T,1,1,1,1,4,4,64,16,1>).  Multi-hundred-megabyte fp16 tensors, 2 bytes in + 2 bytes out per element: it has to
// run at HBM speed, which a thread-per-output gather (16 predicated global loads per output) does not.
//
// Structure: persistent CTAs walk tiles of TW x TH output pixels x one 128-byte channel chunk (64 fp16 / 32 fp32
// channels).  One elected thread keeps a 2-deep ring of TMA box loads in flight per CTA (two CTAs per SM) ((TW+3) x (TH+3) pixels, out-of-bounds
// rows/columns zero-filled by the TMA unit = the op's zero padding), so the loads of the next tiles overlap the
// arithmetic of the current one.  256 threads = 8 column pairs x 8 channel vectors x 4 row groups; a thread walks its rows
// with a rotating set of FH accumulators per column, so every staged row is read from shared memory once (5 16-byte
// loads for 2 output columns) and every output is written with one 16-byte store.
#include "gt_common.cuh"
#include "gt_sm100.cuh"
#include "hot_act.cuh"

using namespace sm100;

namespace {

constexpr int TW = 16, TH = 16, FW = 4, FH = 4, STAGES = 2;
constexpr int BOX_W = TW + FW - 1, BOX_H = TH + FH - 1;
constexpr int STAGE_BYTES = BOX_W * BOX_H * 128;
constexpr int RG = 4, ROWS_PER_THREAD = TH / RG;   // 4 row groups of 4 output rows
constexpr int CPT = 2;                             // output columns per thread

struct FirParams {
    void* y;
    long long ys_n, ys_h, ys_w;
    int N, OH, OW, chunks, tiles_x, tiles_y;
    int padx0, pady0;
    const float* f;        // device taps [FH,FW] with element strides fs_h / fs_w
    long long fs_h, fs_w;
    int flip;
    float gain;
};

__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

// The kernel is issue-bound before it is HBM-bound (ncu, profiles/: 205 warp instructions per 16-byte output vector in
// the first version), so the arithmetic is organised to minimise instructions: every thread owns 2 adjacent output
// columns x 4 output rows x one 16-byte channel vector; per staged row it reads 5 vectors once, converts them to fp32
// once, and -- when the 4x4 filter is an outer product, which every StyleGAN2 resampling filter is
// (upfirdn2d.setup_filter([1,3,3,1])) -- applies the horizontal taps first and feeds the result to the rotating
// vertical accumulators, all with packed fp32x2 FMAs.  A filter that is not rank-1 takes the general 16-tap path.
template <class T>
__global__ void __launch_bounds__(256, 2) upfirdn2d_tma_kernel(const __grid_constant__ CUtensorMap tmX, const FirParams p, const int total_tiles) {
    typedef hot::Lanes<T> L;
    constexpr int NP = L::NP;                      // float2 pairs per 16-byte vector (4 fp16 / 2 fp32)
    constexpr int VEC = Vec16<T>::N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    __shared__ float sf[FH * FW];
    __shared__ float sgx[FW], sgy[FH];
    __shared__ int s_sep;
    if (threadIdx.x < FH * FW) {   // correlation taps: g = f if flip else f reversed, gain folded in (OPS/upfirdn2d.py:196-199)
        const int ky = threadIdx.x / FW, kx = threadIdx.x - ky * FW;
        const int sy = p.flip ? ky : FH - 1 - ky, sx = p.flip ? kx : FW - 1 - kx;
        sf[threadIdx.x] = p.f[sy * p.fs_h + sx * p.fs_w] * p.gain;
    }

    const int tid = threadIdx.x;
    const int cv = tid & 7, cp = (tid >> 3) & 7, rg = tid >> 6;
    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        // rank-1 test: g[ky][kx] == gy[ky] * gx[kx] with gx = pivot row, gy = pivot column / pivot
        int pr = 0, pc = 0;
        float best = 0.f;
        for (int i = 0; i < FH * FW; i++)
            if (fabsf(sf[i]) > best) {
                best = fabsf(sf[i]);
                pr = i / FW;
                pc = i % FW;
            }
        int sep = best > 0.f;
        if (sep) {
            for (int kx = 0; kx < FW; kx++) sgx[kx] = sf[pr * FW + kx];
            for (int ky = 0; ky < FH; ky++) sgy[ky] = sf[ky * FW + pc] / sf[pr * FW + pc];
            for (int i = 0; i < FH * FW; i++)
                if (fabsf(sf[i] - sgy[i / FW] * sgx[i % FW]) > 1e-7f * best) sep = 0;
        }
        s_sep = sep;
    }

    auto issue = [&](int tile, int stage) {
        int t = tile;
        const int ch = t % p.chunks;
        t /= p.chunks;
        const int tx = t % p.tiles_x;
        t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int n = t / p.tiles_y;
        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        tma_load_4d(smem + stage * STAGE_BYTES, &tmX, &full[stage], ch * (128 / (int)sizeof(T)), tx * TW - p.padx0, ty * TH - p.pady0, n);
    };

    const int first = blockIdx.x, step = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            const int tile = first + s * step;
            if (tile < total_tiles) issue(tile, s);
        }
    }
    __syncthreads();
    const bool sep = s_sep != 0;
    float gx[FW], gy[FH];
#pragma unroll
    for (int k = 0; k < FW; k++) gx[k] = sgx[k];
#pragma unroll
    for (int k = 0; k < FH; k++) gy[k] = sgy[k];

    int it = 0;
    for (int tile = first; tile < total_tiles; tile += step, it++) {
        const int stage = it % STAGES;
        mbar_wait(&full[stage], (uint32_t)((it / STAGES) & 1));
        int t = tile;
        const int ch = t % p.chunks;
        t /= p.chunks;
        const int tx = t % p.tiles_x;
        t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int n = t / p.tiles_y;
        const int ox = tx * TW + cp * CPT;
        const int oy0 = ty * TH + rg * ROWS_PER_THREAD;
        const uint32_t sp = smem_u32(smem) + stage * STAGE_BYTES + ((rg * ROWS_PER_THREAD) * BOX_W + cp * CPT) * 128 + cv * 16;
        T* yp = (T*)p.y + (long long)n * p.ys_n + (long long)ox * p.ys_w + (long long)ch * (128 / (int)sizeof(T)) + cv * VEC;

        float2 acc[FH][CPT][NP];
#pragma unroll
        for (int a = 0; a < FH; a++)
#pragma unroll
            for (int c = 0; c < CPT; c++)
#pragma unroll
                for (int k = 0; k < NP; k++) acc[a][c][k] = splat(0.f);
        // staged rows tr = 0 .. ROWS_PER_THREAD + FH - 2 of this thread's strip; output row o = tr - ky
#pragma unroll
        for (int tr = 0; tr < ROWS_PER_THREAD + FH - 1; tr++) {
            float2 v[CPT + FW - 1][NP];
#pragma unroll
            for (int j = 0; j < CPT + FW - 1; j++) {
                Vec16<T> raw;
                *reinterpret_cast<uint4*>(raw.v) = lds128(sp + (tr * BOX_W + j) * 128);
#pragma unroll
                for (int k = 0; k < NP; k++) v[j][k] = L::get(raw, k);
            }
            if (sep) {
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    float2 h[NP];
#pragma unroll
                    for (int k = 0; k < NP; k++) {
                        h[k] = __fmul2_rn(v[c][k], splat(gx[0]));
#pragma unroll
                        for (int kx = 1; kx < FW; kx++) h[k] = __ffma2_rn(v[c + kx][k], splat(gx[kx]), h[k]);
                    }
#pragma unroll
                    for (int ky = 0; ky < FH; ky++) {
                        if (tr - ky >= 0 && tr - ky < ROWS_PER_THREAD) {
#pragma unroll
                            for (int k = 0; k < NP; k++) acc[(tr - ky) % FH][c][k] = __ffma2_rn(h[k], splat(gy[ky]), acc[(tr - ky) % FH][c][k]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int ky = 0; ky < FH; ky++) {
                    if (tr - ky >= 0 && tr - ky < ROWS_PER_THREAD) {
#pragma unroll
                        for (int c = 0; c < CPT; c++)
#pragma unroll
                            for (int kx = 0; kx < FW; kx++) {
                                const float w = sf[ky * FW + kx];
#pragma unroll
                                for (int k = 0; k < NP; k++) acc[(tr - ky) % FH][c][k] = __ffma2_rn(v[c + kx][k], splat(w), acc[(tr - ky) % FH][c][k]);
                            }
                    }
                }
            }
            const int o = tr - (FH - 1);
            if (o >= 0) {
                const int oy = oy0 + o;
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    Vec16<T> out;
#pragma unroll
                    for (int k = 0; k < NP; k++) {
                        L::set(out, k, acc[o % FH][c][k]);
                        acc[o % FH][c][k] = splat(0.f);
                    }
                    if (ox + c < p.OW && oy < p.OH) st16_stream(yp + (long long)oy * p.ys_h + (long long)c * p.ys_w, out);
                }
            }
        }
        __syncthreads();                       // every thread is done reading this stage
        if (tid == 0) {
            const int next = tile + STAGES * step;
            if (next < total_tiles) issue(next, stage);
        }
    }
}

template <class T>
int launch_tma(const CUtensorMap& tm, const FirParams& p, int total_tiles, cudaStream_t st) {
    constexpr int SMEM = STAGES * STAGE_BYTES + STAGES * 8 + 128;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(upfirdn2d_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) {
            gt_set_error("gt_upfirdn2d(tma): cannot reserve %d bytes of shared memory: %s", SMEM, cudaGetErrorString(e));
            return GT_ERR_CUDA;
        }
        configured = true;
    }
    int grid = gt_num_sms() * 2;        // two CTAs per SM (2 x 92 KB of staging): 16 warps hide the shared-memory / FMA latencies
    if (grid > total_tiles) grid = total_tiles;
    upfirdn2d_tma_kernel<T><<<grid, 256, SMEM, st>>>(tm, p, total_tiles);
    GT_CUDA_LAUNCH_CHECK("gt_upfirdn2d(tma)");
    return GT_OK;
}

}  // namespace

// Returns GT_OK when the TMA kernel took the call, -1 when the call is outside its coverage (the caller then uses the
// generic kernels of upfirdn2d.cu), another code on error.  Coverage: 4x4 taps, up = down = 1, fp16 / fp32, channel
// stride 1, C a multiple of one 128-byte chunk, 16-byte aligned pointers and strides.
int gt_upfirdn2d_try_tma(const void* x, const float* f, long long fs_h, long long fs_w, int flip, float gain, void* y, int dtype, int N, int C, int H, int W, long long xs_n, long long xs_h,
                         long long xs_w, int OH, int OW, long long ys_n, long long ys_h, long long ys_w, int padx0, int pady0, cudaStream_t st) {
    const int esz = dtype == GT_F16 ? 2 : 4;
    const int chunk = 128 / esz;
    if ((dtype != GT_F16 && dtype != GT_F32) || C % chunk != 0) return -1;
    if ((((uintptr_t)x) & 15) || (((uintptr_t)y) & 15)) return -1;
    if ((xs_w * esz) % 16 || (xs_h * esz) % 16 || (xs_n * esz) % 16 || (ys_w * esz) % 16 || (ys_h * esz) % 16 || (ys_n * esz) % 16) return -1;
    gt_encode_tiled_fn encode = gt_get_encode_tiled();
    if (!encode) return -1;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)xs_w * esz, (cuuint64_t)xs_h * esz, (cuuint64_t)xs_n * esz};
    cuuint32_t box[4] = {(cuuint32_t)chunk, BOX_W, BOX_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, dtype == GT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(x), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -1;
    FirParams p;
    memset(&p, 0, sizeof(p));
    p.y = y;
    p.ys_n = ys_n;
    p.ys_h = ys_h;
    p.ys_w = ys_w;
    p.N = N;
    p.OH = OH;
    p.OW = OW;
    p.chunks = C / chunk;
    p.tiles_x = (OW + TW - 1) / TW;
    p.tiles_y = (OH + TH - 1) / TH;
    p.padx0 = padx0;
    p.pady0 = pady0;
    p.f = f;
    p.fs_h = fs_h;
    p.fs_w = fs_w;
    p.flip = flip;
    p.gain = gain;
    const long long total = (long long)N * p.tiles_y * p.tiles_x * p.chunks;
    if (total <= 0 || total >= (1ll << 31)) return -1;
    return dtype == GT_F16 ? launch_tma<__half>(tm, p, (int)total, st) : launch_tma<float>(tm, p, (int)total, st);
}
