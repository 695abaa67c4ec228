// trajectories.cu -- expert-demonstration rollouts and their statistics on the device.
//
//   reference: trajectory.py:52-128 (generate_trajectory / generate_trajectories draw
//   `np.random.choice` over a DENSE row p_transition[s, :, a] per step -- O(S) per step, and the row
//   does not exist for large worlds) and maxent.py:15-60 (feature / start-state statistics).
//
// One thread per trajectory; a step is two inverse-cdf draws (action from the policy row, successor
// from the ELL row) from a counter-based generator, Philox4x32-10 keyed by the seed with the
// counter (step, trajectory): every trajectory is a pure function of (seed, index), independent of
// launch shape -- restated on the CPU by the test suite and compared bit for bit.  The
// selection rule is numpy's: normalised running sum, first entry whose cumulative value exceeds u
// (`cdf.searchsorted(u, side='right')`), so zero-probability entries are never chosen.
#include <stdint.h>

#include "common.cuh"
#include "host_util.h"

namespace irlb200 {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ void draw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// 53-bit uniform in [0, 1) from two words (numpy's random_sample recipe)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__global__ void sample_trajectories_kernel(const int S, const int A, const int K, const int32_t *__restrict__ succ_idx,
                                           const double *__restrict__ succ_p, const double *__restrict__ policy,
                                           const double *__restrict__ start_cdf, const uint8_t *__restrict__ term,
                                           const int n_traj, const int max_len, const uint32_t seed_lo,
                                           const uint32_t seed_hi, int32_t *states, int32_t *actions, int32_t *lengths,
                                           double *visit_counts, double *start_counts, int32_t *n_truncated) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_traj) return;
    const Philox rng{seed_lo, seed_hi};
    uint32_t w[4];
    // start state: inverse cdf of the start distribution (trajectory.py:120-124); counter step = 2^32 - 1
    rng.draw(0xffffffffu, 0xffffffffu, (uint32_t)i, 0u, w);
    const double u0 = u53(w[0], w[1]) * start_cdf[S - 1];
    int lo = 0, hi = S - 1;                                 // first s with start_cdf[s] > u0
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (start_cdf[mid] > u0) hi = mid; else lo = mid + 1;
    }
    int s = lo;
    if (start_counts) atomicAdd(start_counts + s, 1.0);
    const size_t row = (size_t)i * ((size_t)max_len + 1);
    int len = 0;
    for (;;) {
        if (states) states[row + len] = s;
        if (visit_counts) atomicAdd(visit_counts + s, 1.0);            // Trajectory.states(): every state_from + the last state_to
        if (term[s]) break;                                            // `while state not in final`   :74
        if (len >= max_len) { atomicAdd(n_truncated, 1); break; }
        rng.draw((uint32_t)len, 0u, (uint32_t)i, 0u, w);
        // action ~ policy[s, :]                                         stochastic_policy_adapter :169
        double tot = 0.0;
        for (int a = 0; a < A; ++a) tot += policy[(size_t)s * A + a];
        const double ua = u53(w[0], w[1]) * tot;
        int act = A - 1;
        double cum = 0.0;
        for (int a = 0; a < A; ++a) {
            cum += policy[(size_t)s * A + a];
            if (cum > ua) { act = a; break; }
        }
        // successor ~ p_transition[s, :, act]                            :76-79
        double ptot = 0.0;
        for (int j = 0; j < K; ++j) ptot += succ_p[((size_t)act * K + j) * S + s];
        const double us = u53(w[2], w[3]) * ptot;
        int nxt = s;
        cum = 0.0;
        for (int j = 0; j < K; ++j) {
            const double pj = succ_p[((size_t)act * K + j) * S + s];
            cum += pj;
            if (pj > 0.0) nxt = succ_idx[(size_t)j * S + s];             // fallback: the last non-zero entry
            if (cum > us) break;
        }
        if (actions) actions[(size_t)i * max_len + len] = act;
        s = nxt;
        ++len;
    }
    lengths[i] = len;
}

}  // namespace irlb200

using namespace irlb200;

extern "C" int irlb200_sample_trajectories(const irlb200_tables *t, const double *policy, const double *start_cdf,
                                           const uint8_t *terminal_mask, int n_traj, int max_len, uint64_t seed,
                                           int32_t *states, int32_t *actions, int32_t *lengths,
                                           double *visit_counts, double *start_counts, int32_t *n_truncated,
                                           void *stream) {
    if (!t || !t->succ_idx || !t->succ_p || t->S <= 0 || t->A <= 0 || t->Ks <= 0)
        return fail(IRLB200_EINVAL, "sample_trajectories: successor table missing");
    if (!policy || !start_cdf || !terminal_mask || !lengths || !n_truncated || n_traj <= 0 || max_len <= 0)
        return fail(IRLB200_EINVAL, "sample_trajectories: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    const int threads = 64;
    sample_trajectories_kernel<<<(n_traj + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
        t->S, t->A, t->Ks, t->succ_idx, t->succ_p, policy, start_cdf, terminal_mask, n_traj, max_len,
        (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), states, actions, lengths, visit_counts, start_counts,
        n_truncated);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "sample_trajectories_kernel");
    return IRLB200_OK;
}
