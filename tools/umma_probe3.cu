// umma_probe3.cu -- tcgen05.mma rate probe with a fully unrolled, constant-folded issue loop (probe 2 showed that a generic issue loop
// costs ~104 clk per MMA in the single issuing thread, hiding the hardware rate).  Patterns are compile-time: every shared-memory
// descriptor is `base + constant`, so the issuing thread spends a handful of instructions per MMA.
// M=128, K=16, kind::f16, SS mode, cta_group::1, dense K-major SWIZZLE_128B tiles in zeroed shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gan_track_b200/csrc -o tools/umma_probe3 tools/umma_probe3.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gt_sm100.cuh"

using namespace sm100;

static const uint32_t SMEM = 200 * 1024;
constexpr uint32_t A0 = 0, B0 = 64 * 1024;

__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc)
        : "memory");
}

// PAT 0: k-loop on one accumulator (16 MMAs: 4 slabs x 4 k)             [GEMM inner loop]
// PAT 1: conv tile: 9 taps x MT sub-tiles x 4 k, MT = 256 / N accumulators, tap-shifted A, B slab per tap   [current halo kernel]
// PAT 2: row streaming: 3 dx x 4 k, one accumulator window of N columns, A shifted by dx pixels, B slab per dx
// PAT 3: row streaming, two strips interleaved per MMA (two independent accumulator windows)
template <int N, int PAT>
__global__ void __launch_bounds__(128, 1) probe_kernel(int iters, long long* out, uint32_t smem_bytes) {
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
    const uint32_t tm = tmem_slot;
    if (warp == 0 && lane == 0) {
        const uint32_t s0 = smem_u32(smem);
        const uint64_t dA = umma_smem_desc(s0 + A0, 0, 1024), dB = umma_smem_desc(s0 + B0, 0, 1024), dAp = umma_smem_desc(s0 + A0, 0, 2304);      // address field is the low 14 bits (>> 4)
        constexpr uint32_t idesc = umma_idesc(128, N, 0, 0, 0);
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            if (PAT == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) mma(tm, dA + (uint64_t)(((i % 4) * 32) >> 4), dB + (uint64_t)(((i / 4) * N * 128 + (i % 4) * 32) >> 4), idesc);
            } else if (PAT == 1) {
                constexpr int MT = 256 / N;
#pragma unroll
                for (int t = 0; t < 9; t++)
#pragma unroll
                    for (int j = 0; j < MT; j++)
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            mma(tm + j * N, dA + (uint64_t)(((t / 3) * 4352 + (t % 3) * 128 + j * 1024 + k * 32) >> 4),
                                dB + (uint64_t)(((t % (N == 256 ? 4 : (N == 128 ? 8 : 9))) * N * 128 + k * 32) >> 4), idesc);
            } else if (PAT == 4 || PAT == 5 || PAT == 6 || PAT == 7) {
                // k-loop on one accumulator with the A start shifted by 128 / 256 / 512 / 640 bytes (a whole number of 128-byte rows)
                constexpr uint32_t sh = PAT == 4 ? 128 : (PAT == 5 ? 256 : (PAT == 6 ? 512 : 640));
#pragma unroll
                for (int i = 0; i < 16; i++) mma(tm, dA + (uint64_t)((sh + (i % 4) * 32) >> 4), dB + (uint64_t)(((i / 4) * N * 128 + (i % 4) * 32) >> 4), idesc);
            } else if (PAT == 8) {
                // k-loop, aligned start, 8-row groups 4352 bytes apart (halo row pitch of a 34-pixel box)
#pragma unroll
                for (int i = 0; i < 16; i++) mma(tm, dAp + (uint64_t)(((i % 4) * 32) >> 4), dB + (uint64_t)(((i / 4) * N * 128 + (i % 4) * 32) >> 4), idesc);
            } else if (PAT == 2) {
#pragma unroll
                for (int dx = 0; dx < 3; dx++)
#pragma unroll
                    for (int k = 0; k < 4; k++) mma(tm, dA + (uint64_t)((dx * 128 + k * 32) >> 4), dB + (uint64_t)((dx * N * 128 + k * 32) >> 4), idesc);
            } else {
#pragma unroll
                for (int dx = 0; dx < 3; dx++)
#pragma unroll
                    for (int k = 0; k < 4; k++)
#pragma unroll
                        for (int s = 0; s < 2; s++)
                            mma(tm + s * 256, dA + (uint64_t)((s * 32768 + dx * 128 + k * 32) >> 4), dB + (uint64_t)((dx * N * 128 + k * 32) >> 4), idesc);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tm, 512);
    }
}

template <int N, int PAT>
static void run(const char* name, int per_iter, long long* d_out, int sms) {
    cudaFuncSetAttribute(probe_kernel<N, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM + 1024);
    const int iters = 4096 / per_iter * 4;
    double cyc = 0;
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(d_out, 0, sizeof(long long) * sms);
        probe_kernel<N, PAT><<<sms, 128, SMEM + 1024>>>(iters, d_out, SMEM);
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
        cyc = sum / sms / iters / per_iter;
    }
    printf("  %-58s %5d %9.1f %9.1f %8.2f\n", name, N, cyc, N / 2.0, (N / 2.0) / cyc);
    fflush(stdout);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d_out;
    cudaMalloc(&d_out, sizeof(long long) * sms);
    printf("# unrolled issue loop; clk per M128 x N x K16 MMA, array floor N/2, fraction of the array rate; %d SMs busy\n", sms);
    printf("# %-58s %5s %9s %9s %8s\n", "pattern", "N", "clk/MMA", "floor", "of array");
    run<64, 0>("k-loop, one accumulator", 16, d_out, sms);
    run<128, 0>("k-loop, one accumulator", 16, d_out, sms);
    run<192, 0>("k-loop, one accumulator", 16, d_out, sms);
    run<256, 0>("k-loop, one accumulator", 16, d_out, sms);
    run<64, 1>("conv tile 9 taps x 4 sub-tiles x 4 k (halo kernel)", 144, d_out, sms);
    run<128, 1>("conv tile 9 taps x 2 sub-tiles x 4 k (halo kernel)", 72, d_out, sms);
    run<256, 1>("conv tile 9 taps x 1 sub-tile x 4 k (halo kernel)", 36, d_out, sms);
    run<64, 4>("k-loop, A start + 128 B", 16, d_out, sms);
    run<64, 5>("k-loop, A start + 256 B", 16, d_out, sms);
    run<64, 6>("k-loop, A start + 512 B", 16, d_out, sms);
    run<64, 7>("k-loop, A start + 640 B", 16, d_out, sms);
    run<64, 8>("k-loop, aligned start, SBO 2304 (18-pixel halo pitch)", 16, d_out, sms);
    run<128, 4>("k-loop, A start + 128 B", 16, d_out, sms);
    run<128, 8>("k-loop, aligned start, SBO 2304 (18-pixel halo pitch)", 16, d_out, sms);
    run<192, 4>("k-loop, A start + 128 B", 16, d_out, sms);
    run<256, 4>("k-loop, A start + 128 B", 16, d_out, sms);
    run<64, 2>("row streaming 3 dx x 4 k, one window", 12, d_out, sms);
    run<128, 2>("row streaming 3 dx x 4 k, one window", 12, d_out, sms);
    run<192, 2>("row streaming 3 dx x 4 k, one window", 12, d_out, sms);
    run<256, 2>("row streaming 3 dx x 4 k, one window", 12, d_out, sms);
    run<192, 3>("row streaming, two strips interleaved", 24, d_out, sms);
    run<128, 3>("row streaming, two strips interleaved", 24, d_out, sms);
    cudaFree(d_out);
    return 0;
}
