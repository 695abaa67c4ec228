// batch_args.cuh -- argument blocks of the batched kernels (passed by value) and the CTA shared-memory carve-up.
#pragma once
#include "phases.cuh"

namespace irlb200 {

// ---------------------------------------------------------------------------
// batched argument blocks (passed by value as kernel parameters)
// ---------------------------------------------------------------------------
struct SuccBatch {
    SuccArgs a;                  // pointers of problem 0
    size_t tab_idx_stride;       // elements between the tables of consecutive problems (0: shared)
    size_t tab_p_stride;
    size_t phi_stride, term_stride;   // 0: shared
    int32_t *n_iter, *status;    // [B] or null
    int out_stride;              // ints between the (n_iter,status) of consecutive problems
};

struct SvfBatch {
    SvfArgs a;
    size_t tab_idx_stride, tab_p_stride;
    size_t p0_stride, term_stride, ef_stride;
    int32_t *n_iter, *status;
    int out_stride;
    const int32_t *order;        // [B] or null: block i works on problem order[i] (a permutation; launch-order
                                 // hint, e.g. longest-first from the previous gradient step's sweep counts)
};

// problem handled by this block of a one-CTA-per-problem forward launch
__device__ __forceinline__ size_t svf_problem(const SvfBatch &bt) {
    return bt.order ? (size_t)__ldg(bt.order + blockIdx.x) : (size_t)blockIdx.x;
}

struct StepBatch {
    SuccArgs s;
    SvfArgs f;
    size_t succ_idx_stride, succ_p_stride, pred_idx_stride, pred_p_stride;
    size_t phi_stride, term_stride, p0_stride, ef_stride;
    int32_t *n_iter, *status;    // [B][2]
    double *policy_out;          // [B][S][A] or null
};

__device__ __forceinline__ void offset_succ(SuccArgs &a, const SuccBatch &bt, size_t b) {
    const size_t S = a.S, A = a.A;
    a.idx += b * bt.tab_idx_stride;
    a.p += b * bt.tab_p_stride;
    a.reward += b * S;
    if (a.phi) a.phi += b * bt.phi_stride;
    if (a.term) a.term += b * bt.term_stride;
    if (a.policy) a.policy += b * S * A;
    if (a.policy2) a.policy2 += b * S * A;
    if (a.value) a.value += b * S;
}

__device__ __forceinline__ void offset_svf(SvfArgs &a, const SvfBatch &bt, size_t b) {
    const size_t S = a.S, A = a.A;
    a.idx += b * bt.tab_idx_stride;
    a.p += b * bt.tab_p_stride;
    a.p0 += b * bt.p0_stride;
    a.term += b * bt.term_stride;
    a.policy += b * S * A;
    if (a.w_scratch) a.w_scratch += b * S * (size_t)a.K;
    a.svf += b * S;
    if (a.grad) {
        a.grad += b * S;
        a.e_features += b * bt.ef_stride;
    }
}

// shared-memory carve-up of the CTA topology: [buf0 | buf1 | scratch(32) | flags | extra...]
__device__ __forceinline__ double *carve_cta(CtaTopo &tp, int S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *base = reinterpret_cast<double *>(smem_raw);
    tp.buf0 = base;
    tp.buf1 = base + S;
    tp.scratch = base + 2 * (size_t)S;
    tp.flag = reinterpret_cast<int *>(tp.scratch + 32);
    tp.vseq = 0;
    return tp.scratch + 34;      // first free double after the flags
}

static size_t cta_smem_bytes(int S, int A, bool with_policy) {
    size_t n = 2 * (size_t)S + 34 + (with_policy ? (size_t)S * A : 0);
    return n * sizeof(double);
}


}  // namespace irlb200
