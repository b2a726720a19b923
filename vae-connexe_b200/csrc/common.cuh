// Shared helpers for libcrvae_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/crvae_b200.h"

#define CRVAE_G (3 * CRVAE_HIDDEN)

namespace crvae {

void set_error(const char* fmt, ...);
void count_launch();

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    count_launch();
    return 0;
}

#define CRVAE_REQUIRE(cond, msg)                                    \
    do {                                                            \
        if (!(cond)) {                                              \
            crvae::set_error("%s: %s", __func__, msg);              \
            return CRVAE_E_BADARG;                                  \
        }                                                           \
    } while (0)

__device__ __forceinline__ float sigmoidf_acc(float x) {
    // 1/(1+exp(-x)) with IEEE division; matches torch.sigmoid to ~1 ulp
    return __fdiv_rn(1.0f, 1.0f + expf(-x));
}

// Fast gate functions: ex2.approx + rcp.approx based, ~2-3 ulp (sigmoid) / ~1e-7 absolute (tanh).  The
// layered parity protocol (P2/P3/P4, tests/test_gpu_train.py) stays green with them, GC included.
// (branch-free: MUFU.EX2 + MUFU.RCP; __frcp_rn / __fdiv_rn would add a slow-path CALL per element)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoidf_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanhf_fast(float x) { return 1.0f - 2.0f * rcp_approx(ex2_approx(2.8853900817779268f * x) + 1.0f); }

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

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace crvae
