// Shared helpers for libpk2.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <string>

#include "../../include/pk2.h"

namespace pk2 {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define PK2_CHECK(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            pk2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                           cudaGetErrorString(_e));                                  \
            return 1;                                                                \
        }                                                                            \
    } while (0)

#define PK2_REQUIRE(cond, ...)                                                       \
    do {                                                                             \
        if (!(cond)) {                                                               \
            pk2::set_error(__VA_ARGS__);                                             \
            return 2;                                                                \
        }                                                                            \
    } while (0)

#define PK2_LAUNCHED() (pk2::g_launches.fetch_add(1, std::memory_order_relaxed))

#define PK2_POST_LAUNCH()                                                            \
    do {                                                                             \
        PK2_LAUNCHED();                                                              \
        PK2_CHECK(cudaGetLastError());                                               \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// log(exp(a)+exp(b)) in double, safe for -inf
__device__ __forceinline__ double log_add(double a, double b) {
    if (a < b) { double t = a; a = b; b = t; }
    if (b == -INFINITY) return a;
    return a + log1p(exp(b - a));
}

}  // namespace pk2
