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
    // optional dictionary-coded successor probabilities (slab_flow.cu, op 1 / 2): p[a][j][i] == dict[code[a][j][i]].
    // A grid world's table holds a handful of distinct values, so one byte per entry replaces eight.
    const uint8_t *code;         // [A][K][cnt] or null
    const double *dict;          // [256]
};

// one state's update with the iterate read through `xl(global index)` and the table values through `pr(a, j)`
template <int OP, int A_T, int K_T, class PR, class XL>
__device__ __forceinline__ double slab_update_p(const OverlapArgs &a, PR pr, XL xl, int i, double *q) {
    const int A = A_T > 0 ? A_T : a.A, K = K_T > 0 ? K_T : a.K;
    if (OP == 3) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j)
            acc = fma(__ldg(a.w + (size_t)j * a.cnt + i), xl(__ldg(a.idx + (size_t)j * a.cnt + i)), acc);
        return __ldg(a.c0 + i) + acc;
    }
    const double k1 = (OP == kOpSoftVI) ? a.c1[i] : 0.0;
    return succ_update<OP, A_T>(
        A, K, pr, [&](int j) { return xl(__ldg(a.idx + (size_t)j * a.cnt + i)); }, a.c0[i], k1, a.discount,
        a.vi_mean, q);
}
template <int OP, int A_T, int K_T, class XL>
__device__ __forceinline__ double slab_update(const OverlapArgs &a, XL xl, int i, double *q) {
    const int K = K_T > 0 ? K_T : a.K;
    return slab_update_p<OP, A_T, K_T>(
        a, [&](int aa, int j) { return __ldg(a.p + ((size_t)aa * K + j) * a.cnt + i); }, xl, i, q);
}

template <int OP, int A_T, int K_T>
__device__ __forceinline__ double overlap_update(const OverlapArgs &a, const double *x_in, int i, double *q) {
    return slab_update<OP, A_T, K_T>(a, [&](int g) { return ld_cg(x_in + g); }, i, q);
}

// ---------------------------------------------------------------------------
// "LL" mailbox slots for boundary rows that cross GPUs (slab_flow.cu): a double travels as two 8-byte
// words, each carrying 32 bits of the value and a 32-bit tag (the iterate number + 1).  8-byte stores
// are single-copy atomic over NVLink, so a reader that finds the expected tag in BOTH words has the
// value -- data and "it has arrived" in one one-way flight, no fence and no separate flag.
// ---------------------------------------------------------------------------
struct __align__(16) LLSlot { unsigned long long w[2]; };

__device__ __forceinline__ void ll_write(LLSlot *p, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
    const unsigned long long w0 = (b & 0xffffffffull) | t, w1 = (b >> 32) | t;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_try_read(const LLSlot *p, unsigned tag, double &v) {
    unsigned long long w0, w1;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    if ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) return false;
    v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    return true;
}

}  // namespace irlb200
