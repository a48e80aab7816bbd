// Shared helpers for the marsb200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/marsb200.h"

namespace marsb200 {

// thread-local message returned by marsb200_last_error()
char* last_error_buffer();

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
    snprintf(last_error_buffer(), 512, fmt, a, b, c);
    return code;
}

#define MARS_REQUIRE(cond, msg)                                                        \
    do {                                                                                 \
        if (!(cond)) return ::marsb200::fail(MARSB200_ERR_ARG, "%s: requirement failed: " msg " (%lld, %lld)", __func__, 0, 0); \
    } while (0)

#define MARS_CUDA_OK(expr)                                                             \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return ::marsb200::fail(MARSB200_ERR_CUDA, "%s: CUDA error %lld at line %lld", cudaGetErrorString(_e), (long long)_e, __LINE__); \
    } while (0)

#define MARS_LAUNCH_OK() MARS_CUDA_OK(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Per-device, thread-safe "done once" flags and cached attributes.  Function attributes
// (cudaFuncAttributeMaxDynamicSharedMemorySize) and the SM count belong to a DEVICE, not to the process: a host
// process that drives several GPUs (or several host threads) must configure each device it touches.
constexpr int MARS_MAX_DEVICES = 64;
struct PerDeviceOnce {
    std::atomic<int> done[MARS_MAX_DEVICES];
};
// Runs `fn` (-> cudaError_t) the first time the calling thread's current device meets `flag`; racing threads may both
// run it (the configured attributes are idempotent), later calls cost one relaxed load.
template <typename Fn>
static inline cudaError_t per_device_once(PerDeviceOnce& flag, Fn fn) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MARS_MAX_DEVICES) return fn();
    if (flag.done[dev].load(std::memory_order_acquire)) return cudaSuccess;
    e = fn();
    if (e == cudaSuccess) flag.done[dev].store(1, std::memory_order_release);
    return e;
}
// SM count of the calling thread's current device (cached per device)
static inline cudaError_t device_sm_count(int* out) {
    static std::atomic<int> cached[MARS_MAX_DEVICES];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool ok = dev >= 0 && dev < MARS_MAX_DEVICES;
    int v = ok ? cached[dev].load(std::memory_order_acquire) : 0;
    if (!v) {
        e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (ok) cached[dev].store(v, std::memory_order_release);
    }
    *out = v;
    return cudaSuccess;
}

__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// adaptive-pool bin i over an axis of length L with g bins: [floor(i*L/g), ceil((i+1)*L/g))
__host__ __device__ static inline int bin_start(int i, int L, int g) { return (int)(((int64_t)i * L) / g); }
__host__ __device__ static inline int bin_end(int i, int L, int g) { return (int)((((int64_t)i + 1) * L + g - 1) / g); }
// bins that contain coordinate x: [floor(x*g/L), ceil((x+1)*g/L) - 1]
__host__ __device__ static inline int bin_lo_of(int x, int L, int g) { return (int)(((int64_t)x * g) / L); }
__host__ __device__ static inline int bin_hi_of(int x, int L, int g) { return (int)((((int64_t)x + 1) * g + L - 1) / L) - 1; }

// 32-bit variants for device hot loops (valid while L <= 32768 and g <= 4096: products stay below 2^31)
__device__ __forceinline__ int bin_start32(int i, int L, int g) { return (int)((uint32_t)(i * L) / (uint32_t)g); }
__device__ __forceinline__ int bin_end32(int i, int L, int g) { return (int)((uint32_t)((i + 1) * L + g - 1) / (uint32_t)g); }
__device__ __forceinline__ int bin_lo_of32(int x, int L, int g) { return (int)((uint32_t)(x * g) / (uint32_t)L); }
__device__ __forceinline__ int bin_hi_of32(int x, int L, int g) { return (int)((uint32_t)((x + 1) * g + L - 1) / (uint32_t)L) - 1; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit load that does not allocate in L1 (inputs are read once)
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// SMs a launch on `s` can use (the stream's green-context partition, else the device): stream_sms.cu
int sms_for_stream(cudaStream_t s, int* out);

// internal back ends shared between translation units
int pairwise_popc(const uint32_t* bits, int E, int P, int64_t wpm, int32_t* inter, cudaStream_t s);
// word_begin / word_count: the pixel slice (packed words of every mask) the launch covers, -1 = to the end; accumulate: add to
// `inter` instead of overwriting it (integer atomics: the slices of a mask may be summed in any order)
int pairwise_mma(const uint32_t* bits, int E, int P, int64_t wpm, int32_t* inter, cudaStream_t s, int64_t word_begin = 0,
                 int64_t word_count = -1, bool accumulate = false);
int pairwise_fp4(const uint32_t* bits, int E, int P, int64_t wpm, int32_t* inter, cudaStream_t s, int64_t word_begin = 0,
                 int64_t word_count = -1, bool accumulate = false);

}  // namespace marsb200
