// umma_probe2.cu -- second tcgen05.mma rate probe: a kernel that replays an arbitrary short "program" of MMAs
// (A offset, B offset, accumulator column, N) so that dependence on the accumulator, on operand reuse and on N can be separated.
// All MMAs are M=128, K=16, kind::f16, SS mode, cta_group::1, A/B dense K-major SWIZZLE_128B tiles (SBO 1024) in zeroed shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gan_track_b200/csrc -o tools/umma_probe2 tools/umma_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gt_sm100.cuh"

using namespace sm100;

constexpr int MAXP = 224;
struct Prog {
    int len, iters;
    uint32_t a_off[MAXP], b_off[MAXP];
    uint16_t d_col[MAXP], n[MAXP];
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Prog p, long long* out, uint32_t smem_bytes) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x * 16; i < smem_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0 && lane == 0) {
        const uint32_t s0 = smem_u32(smem);
        const long long t0 = clock64();
        for (int it = 0; it < p.iters; it++) {
#pragma unroll 1
            for (int i = 0; i < p.len; i++)
                umma_f16(tmem_base + p.d_col[i], umma_smem_desc(s0 + p.a_off[i], 0, 1024), umma_smem_desc(s0 + p.b_off[i], 0, 1024),
                         umma_idesc(128, p.n[i], 0, 0, 0), 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static const uint32_t SMEM = 200 * 1024;
static const uint32_t A0 = 0, B0 = 64 * 1024;        // A region 64 KB, B region 136 KB

static double run(const Prog& p, long long* d_out, int sms, double* per_flop_frac) {
    cudaMemset(d_out, 0, sizeof(long long) * sms);
    probe_kernel<<<sms, 128, SMEM + 1024>>>(p, d_out, SMEM);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        exit(1);
    }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < sms; i++) sum += (double)h[i];
    double macs = 0;
    for (int i = 0; i < p.len; i++) macs += 128.0 * p.n[i] * 16;
    const double cyc = sum / sms / p.iters;          // cycles per program pass
    *per_flop_frac = macs / 4096.0 / cyc;            // fraction of the array rate (4096 MAC/clk/SM)
    return cyc / p.len;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM + 1024) != cudaSuccess) {
        printf("smem\n");
        return 1;
    }
    long long* d_out;
    cudaMalloc(&d_out, sizeof(long long) * sms);
    printf("# M128 x N x K16 SS MMAs; clk per MMA and fraction of the array rate (4096 MAC/clk/SM); %d SMs busy\n", sms);
    printf("# %-62s %5s %9s %8s\n", "pattern", "N", "clk/MMA", "of array");
    auto emit = [&](const char* name, int n, const Prog& p) {
        double frac;
        run(p, d_out, sms, &frac);
        const double c = run(p, d_out, sms, &frac);
        printf("  %-62s %5d %9.1f %8.2f\n", name, n, c, frac);
        fflush(stdout);
    };
    for (int n : {64, 128, 192, 256}) {
        const int nacc_max = 512 / n;
        // P1: everything fixed: pure dependent chain on one accumulator
        {
            Prog p = {};
            p.len = 16;
            p.iters = 256;
            for (int i = 0; i < p.len; i++) p.a_off[i] = A0, p.b_off[i] = B0, p.d_col[i] = 0, p.n[i] = (uint16_t)n;
            emit("P1 same A, same B, same D", n, p);
        }
        // P2: same operands, accumulator changes every MMA over R accumulators
        for (int R : {2, 4}) {
            if (R > nacc_max) continue;
            Prog p = {};
            p.len = 16;
            p.iters = 256;
            for (int i = 0; i < p.len; i++) p.a_off[i] = A0, p.b_off[i] = B0, p.d_col[i] = (uint16_t)((i % R) * n), p.n[i] = (uint16_t)n;
            char nm[96];
            snprintf(nm, sizeof nm, "P2 same A, same B, D cycles over %d accumulators every MMA", R);
            emit(nm, n, p);
        }
        // P3: GEMM k-loop: A and B step through the 4 k-slices of one 128-byte row, D fixed
        {
            Prog p = {};
            p.len = 16;
            p.iters = 256;
            for (int i = 0; i < p.len; i++) p.a_off[i] = A0 + (i % 4) * 32, p.b_off[i] = B0 + (i % 4) * 32, p.d_col[i] = 0, p.n[i] = (uint16_t)n;
            emit("P3 k-loop (A, B step 32 B), same D", n, p);
        }
        // P4: P3 + a new B slab every 4 MMAs (weight taps), D fixed
        {
            Prog p = {};
            p.len = 36;
            p.iters = 128;
            const int nsl = 136 * 1024 / (n * 128) < 9 ? 136 * 1024 / (n * 128) : 9;
            for (int i = 0; i < p.len; i++)
                p.a_off[i] = A0 + (i % 4) * 32, p.b_off[i] = B0 + ((i / 4) % nsl) * n * 128 + (i % 4) * 32, p.d_col[i] = 0, p.n[i] = (uint16_t)n;
            emit("P4 k-loop, new B slab every 4 MMAs, same D", n, p);
        }
        // P5: P4 + A start shifted by (tap % 3) pixels and a different 16 KB tile per tap row, D fixed (a convolution tile, one accumulator)
        {
            Prog p = {};
            p.len = 36;
            p.iters = 128;
            const int nsl = 136 * 1024 / (n * 128) < 9 ? 136 * 1024 / (n * 128) : 9;
            for (int i = 0; i < p.len; i++) {
                const int t = i / 4;
                p.a_off[i] = A0 + (t / 3) * 16384 + (t % 3) * 128 + (i % 4) * 32;
                p.b_off[i] = B0 + (t % nsl) * n * 128 + (i % 4) * 32;
                p.d_col[i] = 0;
                p.n[i] = (uint16_t)n;
            }
            emit("P5 conv taps (A shifts, B slabs), same D", n, p);
        }
        // P6: P5 with two independent accumulators interleaved in runs of 4 MMAs (two tiles in flight)
        for (int R : {2, 4}) {
            if (R > nacc_max) continue;
            Prog p = {};
            p.len = 36 * 2;
            p.iters = 64;
            if (R == 4) continue;
            const int nsl = 136 * 1024 / (n * 128) < 9 ? 136 * 1024 / (n * 128) : 9;
            for (int i = 0; i < p.len; i++) {
                const int g = i / 4, which = g % 2, t = g / 2;
                p.a_off[i] = A0 + which * 32768 + (t / 3) * 8192 + (t % 3) * 128 + (i % 4) * 32;
                p.b_off[i] = B0 + (t % nsl) * n * 128 + (i % 4) * 32;
                p.d_col[i] = (uint16_t)(which * n);
                p.n[i] = (uint16_t)n;
            }
            emit("P6 conv taps, 2 accumulators interleaved in runs of 4", n, p);
        }
        // P7: same as P6 but alternating every MMA
        if (2 <= nacc_max) {
            Prog p = {};
            p.len = 36 * 2;
            p.iters = 64;
            const int nsl = 136 * 1024 / (n * 128) < 9 ? 136 * 1024 / (n * 128) : 9;
            for (int i = 0; i < p.len; i++) {
                const int which = i % 2, j = i / 2, t = j / 4;
                p.a_off[i] = A0 + which * 32768 + (t / 3) * 8192 + (t % 3) * 128 + (j % 4) * 32;
                p.b_off[i] = B0 + (t % nsl) * n * 128 + (j % 4) * 32;
                p.d_col[i] = (uint16_t)(which * n);
                p.n[i] = (uint16_t)n;
            }
            emit("P7 conv taps, 2 accumulators alternating every MMA", n, p);
        }
    }
    // P8: sliding 3-block window of the row-streaming convolution: step r writes 64-column blocks (r, r+1, r+2) mod 8 -- N = 192 when
    // contiguous, split 128 + 64 at the wrap; 12 MMAs per step (3 dx x 4 k)
    for (int strips : {1, 2}) {
        Prog p = {};
        p.iters = 64;
        int len = 0;
        const int nblk = strips == 1 ? 8 : 4;                 // 64-column blocks per strip ring
        for (int r = 0; r < nblk; r++) {
            for (int q = 0; q < 12; q++) {
                for (int s = 0; s < strips; s++) {
                    const int dx = q / 4, k = q % 4;
                    const uint32_t a = A0 + s * 32768 + (r % 2) * 16384 + dx * 128 + k * 32;
                    const uint32_t b = B0 + dx * 192 * 128 + k * 32;
                    const int b0 = r % nblk;
                    const int base = s * nblk * 64;
                    if (b0 + 3 <= nblk) {
                        p.a_off[len] = a, p.b_off[len] = b, p.d_col[len] = (uint16_t)(base + b0 * 64), p.n[len] = 192, len++;
                    } else {
                        const int first = nblk - b0;          // blocks before the wrap (1 or 2)
                        p.a_off[len] = a, p.b_off[len] = b, p.d_col[len] = (uint16_t)(base + b0 * 64), p.n[len] = (uint16_t)(first * 64), len++;
                        p.a_off[len] = a, p.b_off[len] = b + first * 64 * 128, p.d_col[len] = (uint16_t)base, p.n[len] = (uint16_t)((3 - first) * 64), len++;
                    }
                    if (len > MAXP - 2) break;
                }
                if (len > MAXP - 2) break;
            }
            if (len > MAXP - 2) break;
        }
        p.len = len;
        char nm[96];
        snprintf(nm, sizeof nm, "P8 row streaming, N=192 sliding window, %d strip(s) interleaved", strips);
        double frac;
        run(p, d_out, sms, &frac);
        const double c = run(p, d_out, sms, &frac);
        printf("  %-62s %5s %9.1f %8.2f   (%d MMAs per pass)\n", nm, "192*", c, frac, len);
    }
    cudaFree(d_out);
    return 0;
}
