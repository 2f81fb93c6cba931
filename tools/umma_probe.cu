// umma_probe.cu -- measures the issue-to-completion rate of tcgen05.mma (kind::f16, SS mode) on B200 for the operand layouts
// the convolution kernels use, so that kernel design follows measured rates instead of a model:
//   * tile shapes M128 x N{64,128,256} (cta_group::1) and M256 x N{64,128,256} (cta_group::2, CTA pair)
//   * A operand as a dense K-major SWIZZLE_128B tile (SBO 1024) vs the halo-staged view (SBO = halo row pitch, start address
//     shifted by whole 128-byte rows = a convolution tap)
//   * B operand rotating over 9 resident weight slabs (what a weight-stationary kernel would issue)
// Every CTA (grid = #SMs so that the chip runs at its loaded clock) issues `iters` rounds of the 9-tap x MT-sub-tile x 4-k-step
// MMA pattern of the convolution inner loop on zeroed shared memory and reports cycles per MMA (clock64 around issue ... commit).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gan_track_b200/csrc -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gt_sm100.cuh"

using namespace sm100;

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
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((unsigned short)3)
                 : "memory");
}

struct ProbeArgs {
    int iters;          // rounds of the 9 x MT x 4 pattern
    int mt;             // sub-tiles per round sharing one weight slab
    int halo;           // 0: dense A tile per sub-tile (SBO 1024, 16 KB apart); 1: halo view (SBO = pitch, sub-tile j at +j*1024)
    int shift;          // 1: tap t starts (t/3) rows and (t%3) pixels into the halo (unaligned start); 0: every tap at offset 0
    int pitch;          // halo row pitch in bytes
    int b_rotate;       // 1: tap t reads weight slab t (nslabs resident slabs); 0: always slab 0
    int bn;
    int a_kb;           // KB of the A region
    int nslabs;         // weight slabs resident in shared memory
};

// dynamic smem: [A: 96 KB][B: 9 * BN * 128]  (zero filled)
template <int CG>
__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs a, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_bytes = (uint32_t)a.a_kb * 1024u, b_bytes = (uint32_t)a.nslabs * (uint32_t)(a.bn / CG) * 128u;
    for (uint32_t i = threadIdx.x * 16; i < a_bytes + b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if (CG == 1) tmem_alloc(&tmem_slot, 512);
        else tmem_alloc2(&tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    long long cycles = 0;
    if (warp == 0 && lane == 0 && rank == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + a_bytes);
        const uint32_t idesc = umma_idesc(128 * CG, a.bn, 0, 0, 0);
        const uint32_t sbo_a = a.halo ? (uint32_t)a.pitch : 1024u;
        const uint32_t slab = (uint32_t)(a.bn / CG) * 128u;
        const long long t0 = clock64();
        for (int it = 0; it < a.iters; it++) {
            for (int t = 0; t < 9; t++) {
                const uint32_t at = a0 + (a.shift ? (uint32_t)(t / 3) * (uint32_t)a.pitch + (uint32_t)(t % 3) * 128u : 0u);
                const uint32_t bt = b0 + (a.b_rotate ? (uint32_t)(t % a.nslabs) * slab : 0u);
                for (int j = 0; j < a.mt; j++) {
                    const uint32_t aj = at + (a.halo ? (uint32_t)j * 1024u : (uint32_t)j * 16384u);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (CG == 1)
                            umma_f16(tmem_base + (uint32_t)(j * a.bn), umma_smem_desc(aj + k * 32, 0, sbo_a), umma_smem_desc(bt + k * 32, 0, 1024), idesc, 1u);
                        else
                            umma_f16_2sm(tmem_base + (uint32_t)(j * a.bn), umma_smem_desc(aj + k * 32, 0, sbo_a), umma_smem_desc(bt + k * 32, 0, 1024), idesc, 1u);
                    }
                }
            }
        }
        if (CG == 1) umma_commit(&bar);
        else umma_commit_2sm(&bar);
        mbar_wait(&bar, 0);
        cycles = clock64() - t0;
        out[blockIdx.x] = cycles;
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 1) tmem_dealloc(tmem_base, 512);
        else tmem_dealloc2(tmem_base, 512);
    }
}

template <int CG>
static double run(const ProbeArgs& a, long long* d_out, int sms) {
    const size_t smem = (size_t)a.a_kb * 1024 + (size_t)a.nslabs * (size_t)(a.bn / CG) * 128 + 1024;
    if (cudaFuncSetAttribute(probe_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        printf("cannot reserve %zu bytes of shared memory\n", smem);
        exit(1);
    }
    cudaMemset(d_out, 0, sizeof(long long) * sms);
    if (CG == 1) {
        probe_kernel<1><<<sms, 128, smem>>>(a, d_out);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms & ~1);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, probe_kernel<2>, a, d_out);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        exit(1);
    }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double sum = 0;
    int n = 0;
    for (int i = 0; i < sms; i++)
        if (h[i] > 0) {
            sum += (double)h[i];
            n++;
        }
    const double mmas = (double)a.iters * 9 * a.mt * 4;
    return sum / n / mmas;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d_out;
    cudaMalloc(&d_out, sizeof(long long) * sms);
    printf("# tcgen05.mma kind::f16 SS, K=16 per MMA; cycles per MMA (mean over %d SMs, all SMs busy); array-rate floor = M*N/ (cta_group*256) ... M128: N/2 clk\n", sms);
    printf("# %-6s %-4s %-3s %-28s %10s %10s %8s\n", "cta", "N", "MT", "A layout / B", "clk/MMA", "floor", "ratio");
    struct L {
        const char* name;
        int halo, shift, b_rotate;
    } layouts[] = {
        {"dense A, one B slab", 0, 0, 0},
        {"dense A, 9 B slabs", 0, 0, 1},
        {"halo pitch, taps at 0", 1, 0, 1},
        {"dense SBO, shifted start", 0, 1, 1},
        {"halo pitch, shifted taps", 1, 1, 1},
    };
    for (int cg = 1; cg <= 2; cg++) {
        for (int bn : {64, 128, 256}) {
            const int mt = 256 / bn;
            for (const L& l : layouts) {
                ProbeArgs a;
                a.iters = 64;
                a.mt = mt;
                a.halo = l.halo;
                a.shift = l.shift;
                a.pitch = (8 * mt + 2) * 128;
                a.b_rotate = l.b_rotate;
                a.bn = bn;
                a.a_kb = bn == 64 ? 80 : 48;
                a.nslabs = bn == 64 ? 9 : (bn == 128 ? 9 : 4);
                if (cg == 1 && bn == 128) a.nslabs = 9;      // 48 + 144 KB
                if (!l.halo && l.shift) a.pitch = 0;               // dense tile: shift by whole pixels (128-byte rows) only
                run<1>(a, d_out, sms);      // warm-up
                const double c = cg == 1 ? run<1>(a, d_out, sms) : run<2>(a, d_out, sms);
                const double floor_clk = 128.0 * bn / 256.0;      // per MMA of M=128*cg over cg SMs
                printf("  cg::%d  %-4d %-3d %-28s %10.1f %10.1f %8.2f\n", cg, bn, mt, l.name, c, floor_clk, c / floor_clk);
                fflush(stdout);
            }
        }
    }
    // one-sub-tile patterns (MT = 1) for N = 64 / 128: does sharing the B slab across sub-tiles matter?
    for (int bn : {64, 128}) {
        ProbeArgs a;
        a.iters = 128;
        a.mt = 1;
        a.halo = 1;
        a.shift = 1;
        a.pitch = 10 * 128;
        a.b_rotate = 1;
        a.bn = bn;
        a.a_kb = 48;
        a.nslabs = 9;
        const double c = run<1>(a, d_out, sms);
        printf("  cg::1  %-4d %-3d %-28s %10.1f %10.1f %8.2f\n", bn, 1, "halo pitch 10 px, shifted", c, 128.0 * bn / 256.0, c / (128.0 * bn / 256.0));
    }
    // row-streaming form of a 64->64 3x3 convolution: A = one input row strip of 128 pixels (dense, start shifted by dx pixels),
    // B = the three dy taps of one dx stacked along N (192 rows) -> every MMA feeds three output rows
    printf("# row-streaming patterns (A dense 128-pixel strip, shifted start; B = N rows of resident weights)\n");
    for (int bn : {64, 128, 192, 256}) {
        ProbeArgs a;
        a.iters = 128;
        a.mt = 1;
        a.halo = 0;
        a.shift = 1;
        a.pitch = 0;
        a.b_rotate = 1;
        a.bn = bn;
        a.a_kb = 48;
        a.nslabs = bn <= 128 ? 9 : (bn == 192 ? 6 : 4);
        run<1>(a, d_out, sms);
        const double c = run<1>(a, d_out, sms);
        printf("  cg::1  %-4d %-3d %-28s %10.1f %10.1f %8.2f\n", bn, 1, "strip A shifted, N rows of B", c, 128.0 * bn / 256.0, c / (128.0 * bn / 256.0));
    }
    cudaFree(d_out);
    return 0;
}
