// stream_bulk.cuh -- skeleton of the HBM-streaming element-wise kernels (bias_act, modulation, demodulation+activation):
// persistent CTAs move CONTIGUOUS 16 KB chunks global -> shared -> global with the bulk asynchronous copy engine
// (cp.async.bulk + mbarrier complete_tx on the way in, cp.async.bulk bulk_group on the way out), so that each SM keeps
// S-2 chunk loads per input stream and up to two chunk stores in flight without spending registers or LSU issue slots
// on them.  The arithmetic runs IN PLACE on the staged chunk of input 0 (16-byte shared-memory vectors, conflict-free),
// which is then stored from the same buffer.  No CTA-wide barrier in the loop.
//
//   ring:   stage(it) = it % S;   load(it) -> [full mbarrier] -> compute in place -> fence.proxy.async -> [done mbarrier]
//           -> driver lane: store(it), commit, wait until store(it-1) has drained its buffer, refill that stage with load(it-1+S)
//
// Thread t of the CTA handles vectors t, t + NT, t + 2 NT, ... of every chunk; because the chunk size is a multiple of
// NT * 16 bytes, a thread's vectors always sit at the same offsets modulo NT*VEC elements -- which is what lets the
// backward kernels keep per-channel partial sums in registers when the channel count divides NT * VEC.
#pragma once
#include "gt_common.cuh"
#include "gt_sm100.cuh"

namespace streamk {

constexpr int NT = 256;                 // threads per CTA

// NIN input streams, S ring stages, CH_BYTES bytes per chunk per stream (a multiple of NT * 16)
template <int NIN, int S, int CH_BYTES> struct Smem {
    static constexpr int DATA = NIN * S * CH_BYTES;
    static constexpr int SIDE = 512;                       // bytes per stage of the optional side stream (e.g. per-pixel noise)
    static constexpr int TOTAL = DATA + S * SIDE + 2 * S * 8 + 128;   // + side + barriers + slack for manual 128-byte alignment
};

// f(e0, v0, v1, v2, live, side_saddr, vo): e0 = element index of the vector (relative to in0); v0 is updated in place
// (it becomes the output vector); live = false on lanes past the end of a short last chunk; side_saddr = shared address
// of this chunk's side-stream bytes; vo = byte offset of the vector inside the chunk.
// Launch with NTHREADS = NT + 32 threads: warps 0..7 are consumers (thread t < NT owns vectors t, t + NT, ...), warp 8 is
// the copy-engine driver (one lane).  Consumers never wait on a CTA-wide barrier: per stage there is a `full` mbarrier
// (load landed) and a `done` mbarrier (all 8 consumer warps have written their results back and fenced them for the
// async proxy); the driver turns `done` into a bulk store and, once the previous store has drained its buffer, a refill.
constexpr int NTHREADS = NT + 32;

// Optional side stream: `side` (16-byte aligned) supplies side_full bytes (<= 512, a multiple of 16) per full chunk, e.g.
// one value per pixel for a chunk that spans CH_BYTES / (C * sizeof(T)) pixels; it is staged next to the chunk and f
// receives the shared-memory address of its stage.
template <class T, int NIN, int S, int CH_BYTES, class F>
__device__ __forceinline__ void run(const T* in0, const T* in1, const T* in2, T* out, long long nelem, uint8_t* smem_raw, F&& f,
                                    const void* side = nullptr, int side_full = 0) {
    using namespace sm100;
    constexpr int VEC = Vec16<T>::N;
    constexpr int VPT = CH_BYTES / 16 / NT;   // vectors per thread per chunk
    static_assert(CH_BYTES % (NT * 16) == 0, "chunk must be a whole number of vectors per thread");
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    typedef Smem<NIN, S, CH_BYTES> L;
    uint8_t* sside = smem + L::DATA;
    uint64_t* full = (uint64_t*)(smem + L::DATA + S * L::SIDE);
    uint64_t* done = full + S;
    const int tid = threadIdx.x;
    const long long total_bytes = (nelem / VEC) * 16;
    const long long nchunks = (total_bytes + CH_BYTES - 1) / CH_BYTES;
    const long long first = blockIdx.x, step = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&done[s], NT / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= NT) {
        if (tid == NT) {
            const uint8_t* src[3] = {(const uint8_t*)in0, (const uint8_t*)in1, (const uint8_t*)in2};
            auto issue = [&](long long it) {
                const long long c = first + it * step;
                if (c >= nchunks) return;
                const int stage = (int)(it % S);
                const long long off = c * CH_BYTES;
                const uint32_t bytes = (uint32_t)((total_bytes - off) < CH_BYTES ? (total_bytes - off) : CH_BYTES);
                // side bytes of a short last chunk scale with it (both are whole pixels)
                const uint32_t sbytes = side ? (uint32_t)(((long long)bytes * side_full) / CH_BYTES) : 0u;
                mbar_arrive_expect_tx(&full[stage], bytes * NIN + sbytes);
#pragma unroll
                for (int k = 0; k < NIN; k++) bulk_load(smem + (k * S + stage) * CH_BYTES, src[k] + off, bytes, &full[stage]);
                if (sbytes) bulk_load(sside + stage * L::SIDE, (const uint8_t*)side + c * side_full, sbytes, &full[stage]);
            };
            for (int it = 0; it < S; it++) issue(it);
            long long it = 0;
            for (long long c = first; c < nchunks; c += step, it++) {
                const int stage = (int)(it % S);
                const long long off = c * CH_BYTES;
                const uint32_t bytes = (uint32_t)((total_bytes - off) < CH_BYTES ? (total_bytes - off) : CH_BYTES);
                mbar_wait(&done[stage], (uint32_t)((it / S) & 1));
                bulk_store((uint8_t*)out + off, smem + stage * CH_BYTES, bytes);
                bulk_commit();
                if (it >= 1) {
                    bulk_wait_read<1>();          // store(it-1) no longer reads its stage
                    issue(it - 1 + S);
                }
            }
            bulk_wait_all<0>();
        }
        return;
    }

    const uint32_t sa = smem_u32(smem);
    long long it = 0;
    for (long long c = first; c < nchunks; c += step, it++) {
        const int stage = (int)(it % S);
        const long long off = c * CH_BYTES;
        const int bytes = (int)((total_bytes - off) < CH_BYTES ? (total_bytes - off) : CH_BYTES);
        mbar_wait(&full[stage], (uint32_t)((it / S) & 1));
        Vec16<T> v0[VPT], v1[VPT], v2[VPT];
#pragma unroll
        for (int j = 0; j < VPT; j++) {
            const int vo = (tid + j * NT) * 16;
            *reinterpret_cast<uint4*>(v0[j].v) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(v1[j].v) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(v2[j].v) = make_uint4(0, 0, 0, 0);
            if (vo < bytes) {
                *reinterpret_cast<uint4*>(v0[j].v) = lds128(sa + stage * CH_BYTES + vo);
                if (NIN > 1) *reinterpret_cast<uint4*>(v1[j].v) = lds128(sa + (1 * S + stage) * CH_BYTES + vo);
                if (NIN > 2) *reinterpret_cast<uint4*>(v2[j].v) = lds128(sa + (2 * S + stage) * CH_BYTES + vo);
            }
        }
#pragma unroll
        for (int j = 0; j < VPT; j++) {
            const int vo = (tid + j * NT) * 16;
            const bool live = vo < bytes;     // f runs on every lane (it may use warp shuffles); dead lanes hold garbage
            f((off + vo) / (long long)sizeof(T), v0[j], v1[j], v2[j], live, sa + (uint32_t)(L::DATA + stage * L::SIDE), vo);
            if (live) sts128(sa + stage * CH_BYTES + vo, *reinterpret_cast<const uint4*>(v0[j].v));
        }
        fence_proxy_async();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&done[stage]);
    }
}

inline int grid_for(long long nelem, int elem_size, int ch_bytes, int ctas_per_sm) {
    const long long nchunks = (nelem * elem_size + ch_bytes - 1) / ch_bytes;
    const long long cap = (long long)gt_num_sms() * ctas_per_sm;
    long long g = nchunks < cap ? nchunks : cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace streamk
