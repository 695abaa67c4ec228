// slab_kernels.cu -- one sweep over a contiguous SLAB of states of one large MDP.
//
// Multi-GPU mode for a single huge grid (BASELINE configs[4]): every
// rank owns the states [lo, lo + cnt) and holds only their table rows (local
// slot-major arrays of stride cnt, GLOBAL neighbour indices).  The iterate is a
// full-length vector in which the rank keeps its own range plus the ghost rows
// it receives from its neighbours between sweeps.  One launch = one sweep:
// the halo exchange and the convergence all-reduce run between launches
// (irl-maxent_b200/slab.py: NCCL send/recv + all_reduce, exact sweep count by
// snapshot-and-replay).  Same per-state arithmetic as phases.cuh.
#include "host_util.h"
#include "phases.cuh"

namespace irlb200 {

struct SlabSweep {
    int lo, cnt;                 // owned range
    int A, K;
    const int32_t *idx;          // [K][cnt]  global indices
    const double *p;             // [A][K][cnt]   (successor sweeps)   or W [K][cnt] (forward sweep)
    const double *c0;            // [cnt] reward (soft-VI / VI) or p0 (forward)
    const double *c1;            // [cnt] phi (soft-VI) or null
    double discount, eps;
    int vi_mean;
    const double *x_in;          // [S_total]
    double *x_out;               // [S_total], only [lo, lo+cnt) is written
    int *vote;                   // [2]: {some |diff| > eps, some diff is NaN}, OR-ed
    double *policy;              // [cnt][A] or null: written from THIS sweep's q and x (last sweep only)
};

template <int OP, int A_T, int K_T>
__global__ void __launch_bounds__(256) slab_succ_sweep_kernel(const SlabSweep a) {
    const int A = A_T > 0 ? A_T : a.A, K = K_T > 0 ? K_T : a.K;
    constexpr int QN = A_T > 0 ? A_T : kMaxDynA;
    bool gt = false, nan = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.cnt; i += gridDim.x * blockDim.x) {
        double q[QN];
        const double k1 = (OP == kOpSoftVI) ? a.c1[i] : 0.0;
        const double x = succ_update<OP, A_T>(
            A, K, [&](int aa, int j) { return __ldg(a.p + ((size_t)aa * K + j) * a.cnt + i); },
            [&](int j) { return __ldg(a.x_in + __ldg(a.idx + (size_t)j * a.cnt + i)); }, a.c0[i], k1,
            a.discount, a.vi_mean, q);
        Vote v;
        v.reset();
        v.add(x, a.x_in[a.lo + i], a.eps);
        gt |= v.gt;
        nan |= v.nan;
        a.x_out[a.lo + i] = x;
        if (a.policy)
            for (int aa = 0; aa < A; ++aa) a.policy[(size_t)i * A + aa] = exp(q[aa] - x);   // maxent.py:341
    }
    if (__syncthreads_or(gt) && threadIdx.x == 0) atomicOr(&a.vote[0], 1);
    if (__syncthreads_or(nan) && threadIdx.x == 0) atomicOr(&a.vote[1], 1);
}

template <int K_T>
__global__ void __launch_bounds__(256) slab_svf_sweep_kernel(const SlabSweep a) {
    const int K = K_T > 0 ? K_T : a.K;
    bool gt = false, nan = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.cnt; i += gridDim.x * blockDim.x) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j)
            acc = fma(__ldg(a.p + (size_t)j * a.cnt + i), __ldg(a.x_in + __ldg(a.idx + (size_t)j * a.cnt + i)), acc);
        const double x = a.c0[i] + acc;
        Vote v;
        v.reset();
        v.add(x, a.x_in[a.lo + i], a.eps);
        gt |= v.gt;
        nan |= v.nan;
        a.x_out[a.lo + i] = x;
    }
    if (__syncthreads_or(gt) && threadIdx.x == 0) atomicOr(&a.vote[0], 1);
    if (__syncthreads_or(nan) && threadIdx.x == 0) atomicOr(&a.vote[1], 1);
}

// W[j][i] = sum_a pred_p[a][j][i] * policy[pred][a], 0 if pred is terminal; policy and the mask are
// full-length, globally indexed (the ghost rows of the policy were exchanged by the caller)
__global__ void slab_weights_kernel(int cnt, int A, int K, const int32_t *idx, const double *p,
                                    const double *policy, const uint8_t *term, double *W) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x)
        for (int j = 0; j < K; ++j) {
            const int pred = idx[(size_t)j * cnt + i];
            double acc = 0.0;
            for (int aa = 0; aa < A; ++aa)
                acc = fma(p[((size_t)aa * K + j) * cnt + i], policy[(size_t)pred * A + aa], acc);
            W[(size_t)j * cnt + i] = term[pred] ? 0.0 : acc;
        }
}

static int slab_blocks(int cnt) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = ((long long)cnt + 255) / 256, cap = (long long)sms * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace irlb200

using namespace irlb200;

extern "C" int irlb200_slab_sweep(int op, int lo, int cnt, int A, int K, const int32_t *idx,
                                  const double *p, const double *c0, const double *c1,
                                  double discount, double eps, int vi_mean, const double *x_in,
                                  double *x_out, int32_t *vote, double *policy, void *stream) {
    if (cnt <= 0 || !idx || !p || !c0 || !x_in || !x_out || !vote) return fail(IRLB200_EINVAL, "slab_sweep: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    SlabSweep a{lo, cnt, A, K, idx, p, c0, c1, discount, eps, vi_mean, x_in, x_out, vote, policy};
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = slab_blocks(cnt);
    const bool fast = (A == 4 && K == 5);
    if (!fast && A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
    switch (op) {
        case 1:     // soft value iteration sweep
            if (!c1) return fail(IRLB200_EINVAL, "slab_sweep: soft-VI needs phi");
            if (fast) slab_succ_sweep_kernel<kOpSoftVI, 4, 5><<<blocks, 256, 0, st>>>(a);
            else slab_succ_sweep_kernel<kOpSoftVI, 0, 0><<<blocks, 256, 0, st>>>(a);
            break;
        case 2:     // value iteration sweep
            if (fast) slab_succ_sweep_kernel<kOpVI, 4, 5><<<blocks, 256, 0, st>>>(a);
            else slab_succ_sweep_kernel<kOpVI, 0, 0><<<blocks, 256, 0, st>>>(a);
            break;
        case 3:     // forward sweep, p = W [K][cnt]
            if (K == 5) slab_svf_sweep_kernel<5><<<blocks, 256, 0, st>>>(a);
            else slab_svf_sweep_kernel<0><<<blocks, 256, 0, st>>>(a);
            break;
        default:
            return fail(IRLB200_EINVAL, "slab_sweep: unknown op");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "slab sweep kernel");
    return IRLB200_OK;
}

extern "C" int irlb200_slab_weights(int cnt, int A, int K, const int32_t *pred_idx, const double *pred_p,
                                    const double *policy, const uint8_t *terminal_mask, double *W,
                                    void *stream) {
    if (cnt <= 0 || !pred_idx || !pred_p || !policy || !terminal_mask || !W) return fail(IRLB200_EINVAL, "slab_weights: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    slab_weights_kernel<<<slab_blocks(cnt), 256, 0, (cudaStream_t)stream>>>(cnt, A, K, pred_idx, pred_p, policy,
                                                                            terminal_mask, W);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "slab_weights_kernel");
    return IRLB200_OK;
}
