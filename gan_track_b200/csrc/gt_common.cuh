// gan_track_b200 -- shared helpers for the sm_100a kernels behind the C ABI (include/gantrack_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

// dtype codes used across the C ABI.
enum { GT_F32 = 0, GT_F16 = 1, GT_F64 = 2 };

// ---- error handling: return code + thread-local message (SURVEY.md section 8b "Errors") -------------------------
#define GT_OK 0
#define GT_ERR_ARG 1
#define GT_ERR_CUDA 2
#define GT_ERR_UNSUPPORTED 3

void gt_set_error(const char* fmt, ...);

#define GT_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            gt_set_error(__VA_ARGS__);        \
            return GT_ERR_ARG;                \
        }                                     \
    } while (0)

#define GT_CUDA_LAUNCH_CHECK(name)                                                        \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            gt_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));    \
            return GT_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

// ---- device properties (cached per process; the library keeps no other global state) ---------------------------
int gt_num_sms();
int gt_stream_variant();   // 0 = bulk-copy staged streaming kernels, 1 = direct vector loads/stores (tuning / A-B)


// ---- division by a run-time constant without the ~40-instruction integer divide --------------------------------
// floor(n / d) for 0 <= n < 2^31 and 1 <= d < 2^31 (Granlund-Montgomery: m = ceil(2^(31+l) / d), l = ceil(log2 d)).
struct FastDiv {
    uint32_t m, sh, d;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t l = 0;
    while ((1ull << l) < d) l++;
    f.m = (uint32_t)(((1ull << (31 + l)) + d - 1) / d);
    f.sh = 31 + l;
    f.d = d;
    return f;
}
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv& f) { return (uint32_t)(((unsigned long long)n * f.m) >> f.sh); }
// n -> (n / d, n % d)
__device__ __forceinline__ uint32_t fd_divmod(uint32_t n, const FastDiv& f, uint32_t& rem) {
    const uint32_t q = fd_div(n, f);
    rem = n - q * f.d;
    return q;
}

// ---- scalar type traits: I/O type -> accumulation type (fp16 I/O computes in fp32, like OPS/bias_act.cu:15-18) --
template <class T> struct Acc { typedef float type; };
template <> struct Acc<double> { typedef double type; };

template <class T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v) { return (typename Acc<T>::type)v; }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }

template <class T> __device__ __forceinline__ T from_acc(typename Acc<T>::type v) { return (T)v; }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half_rn(v); }

// ---- 16-byte vector I/O ---------------------------------------------------------------------------------------
template <class T> struct alignas(16) Vec16 { static constexpr int N = 16 / sizeof(T); T v[16 / sizeof(T)]; };

template <class T> __device__ __forceinline__ Vec16<T> ld16(const T* p) {
    Vec16<T> r;
    *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p);
    return r;
}
// streaming (read-once) load: bypass L1 allocation
template <class T> __device__ __forceinline__ Vec16<T> ld16_stream(const T* p) {
    Vec16<T> r;
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    *reinterpret_cast<uint4*>(r.v) = u;
    return r;
}
template <class T> __device__ __forceinline__ void st16(T* p, const Vec16<T>& r) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(r.v);
}
template <class T> __device__ __forceinline__ void st16_stream(T* p, const Vec16<T>& r) {
    uint4 u = *reinterpret_cast<const uint4*>(r.v);
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// fp16x3 route of the fp32 convolutions (csrc/conv_f16x3.cu): power-of-two scale that maps a tensor whose max |v| has the given fp32
// bit pattern into [2^13, 2^14); zero / non-finite maxima -> 1
__device__ __forceinline__ float gt_scale_from_amax_bits(uint32_t bits) {
    const int e = (int)(bits >> 23) & 0xff;
    if (e == 0 || e == 0xff) return 1.f;
    return __int_as_float((uint32_t)(127 + 13 - (e - 127)) << 23);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
