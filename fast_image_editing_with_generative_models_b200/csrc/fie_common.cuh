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

// x * sigmoid(x) with two MUFU ops (ex2, rcp); the IEEE division it replaces cost ~10 extra issue slots per element.
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// Exact (erf) GELU, 0.5 x (1 + erf(x / sqrt 2)), with erfc from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below
// the fp16 rounding of the result): q = erfc(|z|) = poly5(t) exp(-z^2), t = 1 / (1 + p |z|).  Written on the erfc side
// so the negative tail (1 + erf = q) has no cancellation.  ~14 issue slots and 2 MUFU instead of erff's ~30 + branch.
__device__ __forceinline__ float gelu_erf_f(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
    const float q = pl * t * e;
    return 0.5f * x * (x < 0.0f ? q : 2.0f - q);
}

__device__ __forceinline__ float quick_gelu_f(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }

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
