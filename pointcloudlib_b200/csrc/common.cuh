// common.cuh — shared helpers for the libpcl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pcl_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpcl_b200 is written for sm_100a (B200) only"
#endif

namespace pcl {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char *fmt, ...);

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return PCL_OK;
}

#define PCL_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            pcl::set_error(__VA_ARGS__); \
            return PCL_ERR_INVALID_ARG; \
        }                               \
    } while (0)

// Squared distance with the rounding sequence nvcc emits for the reference's
// (a-b)*(a-b) + (c-d)*(c-d) + (e-f)*(e-f)  (misc/ops.py:165, :317): mul, fma, fma.
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by,
                                         float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Matmul-form squared distance of misc/ops.py:30-51 in the oracle's canonical arithmetic:
// inner = fma chain (first term a plain product), *(-2), + |a|^2, + |b|^2.
__device__ __forceinline__ float sqnorm_c(const float *v, int C) {
    float s = __fmul_rn(v[0], v[0]);
    for (int c = 1; c < C; ++c) s = __fadd_rn(s, __fmul_rn(v[c], v[c]));
    return s;
}
__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float sqdist_mm3(float ax, float ay, float az, float na, float bx,
                                            float by, float bz, float nb) {
    float inner = __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
    float d = __fmul_rn(-2.0f, inner);
    d = __fadd_rn(d, na);
    d = __fadd_rn(d, nb);
    return d;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace pcl
