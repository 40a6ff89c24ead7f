// Shared helpers for the fie_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/fie_b200.h"

namespace fie {

// Thread-local last-error string (the only mutable global state of the library).
void set_error(const char* fmt, ...);
int  check_launch(const char* what);   // returns FIE_OK or FIE_ERR_CUDA after cudaGetLastError()

#define FIE_REQUIRE(cond, ...)                                   \
    do { if (!(cond)) { fie::set_error(__VA_ARGS__); return FIE_ERR_INVALID; } } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

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

}  // namespace fie
