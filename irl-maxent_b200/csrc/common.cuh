// common.cuh -- device helpers shared by the sweep kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/irl_maxent_b200.h"

namespace irlb200 {

constexpr double kNegHuge = -1e200;      // reference: maxent.py:323

// ---------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------
// L2-coherent accesses for iterate vectors that other SMs rewrite every sweep
// (L1 is not coherent across SMs, so these must bypass it).
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(double *p, double v) { __stcg(p, v); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_s32(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_s32(int *p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------
// arithmetic of the reference, restated once
// ---------------------------------------------------------------------------
// maxent.py:260-276: max + log(1 + exp(min - max)), plain log (not log1p).
// np.maximum / np.minimum propagate NaN, fmax/fmin do not -> patch it up.
__device__ __forceinline__ double softmax2(double x1, double x2) {
    double hi = fmax(x1, x2);
    double lo = fmin(x1, x2);
    double r = hi + log(1.0 + exp(lo - hi));
    return (x1 != x1 || x2 != x2) ? (x1 + x2) : r;
}

// np.max semantics (NaN wins)
__device__ __forceinline__ double max_nan(double m, double q) {
    return (q > m || q != q) ? q : m;
}

// exact scale by 2^e (|e| small enough that 2^e is a normal double)
__device__ __forceinline__ double scale_pow2(double x, int e) {
    return x * __longlong_as_double((long long)(1023 + e) << 52);
}

// exponent as returned by frexp: x = m * 2^e with 0.5 <= m < 1
__device__ __forceinline__ int frexp_exponent(double x) {
    int e;
    (void)frexp(x, &e);
    return e;
}

// ---------------------------------------------------------------------------
// block-level reductions (any block size that is a multiple of 32, <= 1024)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = (u > v || u != u) ? u : v;      // NaN wins
    }
    return v;
}

// max over the block; every thread gets the result.  `scratch` >= 32 doubles.
__device__ __forceinline__ double block_max(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();                        // protect scratch from a previous use
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = scratch[0];
    for (int i = 1; i < nw; ++i) {
        double u = scratch[i];
        r = (u > r || u != u) ? u : r;
    }
    return r;
}

// ---------------------------------------------------------------------------
// stop-rule vote shared by every convergence loop
//   reference: `while delta > eps` with delta = max |x' - x| (maxent.py:108,112,
//   326,338; solver.py:40,50): keep going iff some |diff| > eps and no diff is NaN.
// ---------------------------------------------------------------------------
struct Vote {
    bool gt;     // some |diff| > eps on this thread
    bool nan;    // some diff is NaN on this thread
    __device__ __forceinline__ void reset() { gt = false; nan = false; }
    __device__ __forceinline__ void add(double x_new, double x_old, double eps) {
        double diff = fabs(x_new - x_old);
        gt |= (diff > eps);
        nan |= (diff != diff);
    }
};

enum : int { kContinue = -1 };

}  // namespace irlb200
