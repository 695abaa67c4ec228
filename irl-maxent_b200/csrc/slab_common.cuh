// slab_common.cuh -- definitions shared by the multi-GPU slab kernels (slab_persistent.cu, slab_flow.cu):
// the peer-mapped block header, the peer pointer table, system-scope access helpers and the per-state
// update of a slab (same arithmetic as phases.cuh, streamed flavour, global neighbour indices).
#pragma once
#include "phases.cuh"

namespace irlb200 {

constexpr int kMaxRanks = 16;

struct SlabShared {                               // start of every rank's peer-mapped block
    unsigned long long flags[2][kMaxRanks];       // [seq parity][source rank]: (seq + 1) << 8 | votes
    unsigned long long slot[4];                   // local arrivals [19:0] + gt votes [39:20] + nan votes [59:40]
    unsigned long long release[4];                // (seq + 1) << 8 | decision bits {1: gt, 2: nan, 4: abort}
    unsigned long long dflag[2][2];               // overlap kernels: [seq parity][0: from rank-1, 1: from rank+1] = seq + 1
    unsigned int bcount[2][2];                    // overlap kernels: CTAs of a boundary group that finished their share
};
constexpr size_t kSlabHeaderBytes = 4096;          // SlabShared at 0, FlowShared (slab_flow.cu) at 1024

struct SlabPeers {
    SlabShared *shared[kMaxRanks];                // every rank's header (own entry = local pointer)
    double *lo_buf0, *lo_buf1;                    // iterate buffers of rank - 1 (null at the first rank)
    double *hi_buf0, *hi_buf1;                    // iterate buffers of rank + 1 (null at the last rank)
    int rank, world;
    int lo, hi, halo;                             // owned global range and ghost width
    long long timeout_ns;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct OverlapArgs {
    int op;                      // 1 soft-VI, 2 VI, 3 forward
    int lo, cnt, S_total, halo;
    int A, K;
    const int32_t *idx;          // [K][cnt] global indices
    const double *p;             // [A][K][cnt] (op 1, 2) or predecessor rows (op 3)
    const double *c0, *c1;       // reward / p_initial, phi
    const double *policy_in;     // op 3: [S_total][A]
    const uint8_t *term;         // op 3: [S_total]
    double *w;                   // op 3: [K][cnt] scratch
    double discount, eps;
    int max_sweeps, vi_mean;
    double *out, *policy_out;
};

template <int OP, int A_T, int K_T>
__device__ __forceinline__ double overlap_update(const OverlapArgs &a, const double *x_in, int i, double *q) {
    const int A = A_T > 0 ? A_T : a.A, K = K_T > 0 ? K_T : a.K;
    if (OP == 3) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j)
            acc = fma(__ldg(a.w + (size_t)j * a.cnt + i), ld_cg(x_in + __ldg(a.idx + (size_t)j * a.cnt + i)), acc);
        return __ldg(a.c0 + i) + acc;
    }
    const double k1 = (OP == kOpSoftVI) ? a.c1[i] : 0.0;
    return succ_update<OP, A_T>(
        A, K, [&](int aa, int j) { return __ldg(a.p + ((size_t)aa * K + j) * a.cnt + i); },
        [&](int j) { return ld_cg(x_in + __ldg(a.idx + (size_t)j * a.cnt + i)); }, a.c0[i], k1, a.discount,
        a.vi_mean, q);
}

}  // namespace irlb200
