// kernels_cta.cuh -- one CTA (or one warp) per world: generic phases, fused gradient step, hand-tuned gather forward kernel.
#pragma once
#include "batch_args.cuh"

namespace irlb200 {

// ---------------------------------------------------------------------------
// CTA kernels
// ---------------------------------------------------------------------------
template <int OP, int A_T, int K_T, int SPT_T, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) succ_cta_kernel(const SuccBatch bt) {
    CtaTopo tp;
    SuccArgs a = bt.a;
    carve_cta(tp, a.S);
    offset_succ(a, bt, blockIdx.x);
    int *ni = bt.n_iter ? bt.n_iter + (size_t)blockIdx.x * bt.out_stride : nullptr;
    int *st = bt.status ? bt.status + (size_t)blockIdx.x * bt.out_stride : nullptr;
    succ_phase<CtaTopo, OP, A_T, K_T, SPT_T>(tp, a, ni, st);
}

template <int A_T, int K_T, int SPT_T, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) svf_cta_kernel(const SvfBatch bt) {
    CtaTopo tp;
    SvfArgs a = bt.a;
    carve_cta(tp, a.S);
    const size_t b = svf_problem(bt);
    offset_svf(a, bt, b);
    int *ni = bt.n_iter ? bt.n_iter + b * bt.out_stride : nullptr;
    int *st = bt.status ? bt.status + b * bt.out_stride : nullptr;
    svf_phase<CtaTopo, A_T, K_T, SPT_T>(tp, a, ni, st);
}

template <bool CAUSAL, int A_T, int K_T, int SPT_T, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) step_cta_kernel(const StepBatch bt) {
    CtaTopo tp;
    SuccArgs s = bt.s;
    SvfArgs f = bt.f;
    double *pol = carve_cta(tp, s.S);
    const size_t b = blockIdx.x, S = s.S, A = s.A;
    s.idx += b * bt.succ_idx_stride;
    s.p += b * bt.succ_p_stride;
    s.reward += b * S;
    if (s.phi) s.phi += b * bt.phi_stride;
    if (s.term) s.term += b * bt.term_stride;
    s.policy = pol;
    s.policy2 = bt.policy_out ? bt.policy_out + b * S * A : nullptr;
    s.value = nullptr;
    f.idx += b * bt.pred_idx_stride;
    f.p += b * bt.pred_p_stride;
    f.p0 += b * bt.p0_stride;
    f.term += b * bt.term_stride;
    f.policy = pol;
    if (f.w_scratch) f.w_scratch += b * S * (size_t)f.K;
    f.svf += b * S;
    if (f.grad) {
        f.grad += b * S;
        f.e_features += b * bt.ef_stride;
    }
    int *ni = bt.n_iter ? bt.n_iter + 2 * b : nullptr;
    int *st = bt.status ? bt.status + 2 * b : nullptr;
    succ_phase<CtaTopo, CAUSAL ? kOpSoftVI : kOpBackward, A_T, K_T, SPT_T>(tp, s, ni, st);
    svf_phase<CtaTopo, A_T, K_T, SPT_T>(tp, f, ni ? ni + 1 : nullptr, st ? st + 1 : nullptr);
}

// ---------------------------------------------------------------------------
// Warp-per-world fused gradient step for tiny worlds (S <= 32, A = 4, K = 5): BASELINE configs[0..1].
//
// A 5x5 world is one warp of work; with one CTA per world every sweep pays a bar.red round trip
// (~75 cycles) and two shared-memory latencies for nothing.  Here a lane IS a state: the iterate
// lives in one register per lane, a neighbour's value is `__shfl_sync(x, idx)`, the stop rule is
// `__any_sync` -- no barrier, no shared memory -- and a batch packs four worlds per CTA.
// Per-state arithmetic is succ_update / the forward FMA chain, exactly as in the CTA kernels, so
// results are bitwise identical to them.
// ---------------------------------------------------------------------------
// The world of one warp: table rows of lane's state in registers, loaded once.
struct WarpWorld {
    static constexpr int A = 4, K = 5;
    int ix[K], px[K];            // successor / predecessor lanes
    double pr[A][K];             // P[s, succ_j, a]
    double pp[A][K];             // P[pred_j, s, a]
    double c1;                   // phi (causal)
    double p0;                   // start mass
    int is_term;                 // terminal mask of this state
    bool seed;                   // backward pass: zs starts at 1 here (terminal)
    bool act;
};

template <bool CAUSAL>
__device__ __forceinline__ void warp_world_load(WarpWorld &wd, const StepBatch &bt, const size_t b, const int lane) {
    constexpr int A = 4, K = 5;
    SuccArgs s = bt.s;
    SvfArgs f = bt.f;
    const int S = s.S;
    wd.act = lane < S;
    const int me = wd.act ? lane : 0;
    s.idx += b * bt.succ_idx_stride; s.p += b * bt.succ_p_stride;
    if (s.phi) s.phi += b * bt.phi_stride;
    if (s.term) s.term += b * bt.term_stride;
    f.idx += b * bt.pred_idx_stride; f.p += b * bt.pred_p_stride; f.p0 += b * bt.p0_stride; f.term += b * bt.term_stride;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        wd.ix[j] = wd.act ? s.idx[(size_t)j * S + me] : 0;
        wd.px[j] = wd.act ? f.idx[(size_t)j * S + me] : 0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            wd.pr[a][j] = wd.act ? s.p[((size_t)a * K + j) * S + me] : 0.0;
            wd.pp[a][j] = wd.act ? __ldg(f.p + ((size_t)a * K + j) * S + me) : 0.0;
        }
    }
    wd.c1 = (CAUSAL && wd.act) ? s.phi[me] : 0.0;
    wd.p0 = wd.act ? f.p0[me] : 0.0;
    wd.is_term = (wd.act && f.term[me]) ? 1 : 0;
    wd.seed = !CAUSAL && wd.act && s.term[me];
}

// One gradient-step body of a warp's world: policy pass (backward sweeps or soft-VI) on reward `r`
// (one value per lane), forward pass; returns the state-visitation value of the lane's state.
template <bool CAUSAL>
__device__ __forceinline__ double warp_world_step(const WarpWorld &wd, const StepBatch &bt, const double r,
                                                  double (&pol)[4], int &n_pol, int &st_pol, int &n_svf, int &st_svf) {
    constexpr int A = 4, K = 5;
    constexpr unsigned FULL = 0xffffffffu;
    const SuccArgs &s = bt.s;
    const SvfArgs &f = bt.f;
    const bool act = wd.act;
    // ---- policy pass -----------------------------------------------------------------------------
    const double c0 = CAUSAL ? r : exp(r);
    double x = CAUSAL ? kNegHuge : (wd.seed ? 1.0 : 0.0);
    double x_old = x;
    n_pol = 0;
    st_pol = IRLB200_ST_CONVERGED;
    auto gather_update = [&](double xin, double *q) {
        double xv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) xv[j] = __shfl_sync(FULL, xin, wd.ix[j]);
        return succ_update<CAUSAL ? kOpSoftVI : kOpBackward, 4>(
            A, K, [&](int a, int j) { return wd.pr[a][j]; }, [&](int j) { return xv[j]; }, c0, wd.c1, s.discount, 0, q);
    };
    if (CAUSAL) {
        const int limit = s.max_sweeps > 0 ? s.max_sweeps : 0x7fffffff;
        for (;;) {
            const double xn = gather_update(x, nullptr);
            const double diff = fabs(xn - x);
            x_old = x;
            x = xn;
            ++n_pol;
            const bool nan = __any_sync(FULL, act && diff != diff);
            const bool gt = __any_sync(FULL, act && diff > s.eps);
            if (nan) { st_pol = IRLB200_ST_NONFINITE; break; }
            if (!gt) break;
            if (n_pol >= limit) { st_pol = IRLB200_ST_MAXSWEEPS; break; }
        }
    } else {
        double mr = act ? fabs(r) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mr = fmax(mr, __shfl_xor_sync(FULL, mr, o));
        const int R = backward_rescale_period(mr, A);
        for (int t = 0; t < s.n_sweeps; ++t) {
            x_old = x;
            x = gather_update(x, nullptr);
            ++n_pol;
            if (n_pol % R == 0 && n_pol < s.n_sweeps) {
                const double m = warp_max(act ? x : 0.0);
                if (m > 0.0 && m < INFINITY) x = ldexp(x, -frexp_exponent(m));
            }
        }
    }
    {
        double q[A];
        const double xr = gather_update(x_old, q);           // the last sweep's per-action terms, bit for bit
#pragma unroll
        for (int a = 0; a < A; ++a) pol[a] = CAUSAL ? exp(q[a] - x) : q[a] / xr;
        if (!CAUSAL && n_pol == 0) {
#pragma unroll
            for (int a = 0; a < A; ++a) pol[a] = 0.0;
        }
    }
    // ---- forward pass ------------------------------------------------------------------------------
    double w[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const double pa = __shfl_sync(FULL, pol[a], wd.px[j]);              // policy[pred_j, a]
            acc = fma(wd.pp[a][j], pa, acc);
        }
        const int pterm = __shfl_sync(FULL, wd.is_term, wd.px[j]);
        w[j] = (pterm || !act) ? 0.0 : acc;
    }
    double d = 0.0;
    n_svf = 0;
    st_svf = IRLB200_ST_CONVERGED;
    const int limit = f.max_sweeps > 0 ? f.max_sweeps : 0x7fffffff;
    for (;;) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) acc = fma(w[j], __shfl_sync(FULL, d, wd.px[j]), acc);
        const double dn = wd.p0 + acc;
        const double diff = fabs(dn - d);
        d = dn;
        ++n_svf;
        const bool nan = __any_sync(FULL, act && diff != diff);
        const bool gt = __any_sync(FULL, act && diff > f.eps);
        if (nan) { st_svf = IRLB200_ST_NONFINITE; break; }
        if (!gt) break;
        if (n_svf >= limit) { st_svf = IRLB200_ST_MAXSWEEPS; break; }
    }
    return d;
}

template <bool CAUSAL>
__global__ void __launch_bounds__(128) step_warp_kernel(const StepBatch bt, const int B) {
    constexpr int A = 4;
    const int lane = threadIdx.x & 31;
    const size_t b = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= (size_t)B) return;                                   // whole warps leave: no block-level sync below
    const int S = bt.s.S;
    WarpWorld wd;
    warp_world_load<CAUSAL>(wd, bt, b, lane);
    const int me = wd.act ? lane : 0;
    const double r = wd.act ? bt.s.reward[b * S + me] : 0.0;
    double pol[A];
    int n_pol, st_pol, n_svf, st_svf;
    const double d = warp_world_step<CAUSAL>(wd, bt, r, pol, n_pol, st_pol, n_svf, st_svf);
    if (wd.act && bt.policy_out) {
#pragma unroll
        for (int a = 0; a < A; ++a) bt.policy_out[(b * S + me) * A + a] = pol[a];
    }
    if (wd.act) {
        bt.f.svf[b * S + me] = d;
        if (bt.f.grad) bt.f.grad[b * S + me] = bt.f.e_features[b * bt.ef_stride + me] - d;
    }
    if (lane == 0) {
        if (bt.n_iter) { bt.n_iter[2 * b] = n_pol; bt.n_iter[2 * b + 1] = n_svf; }
        if (bt.status) { bt.status[2 * b] = st_pol; bt.status[2 * b + 1] = st_svf; }
    }
}

// ---------------------------------------------------------------------------
// The whole outer loop of `irl` / `irl_causal` (maxent.py:240-252, :436-450) for tiny worlds with
// identity features and a built-in optimizer, in ONE launch: per gradient step the warp runs
// warp_world_step on reward = omega, forms grad = e_features - svf (:248), applies the optimizer's
// update rule with the learning rate the HOST evaluated for that step (the schedule stays a host-side
// Python callable: `lr[k]` is its value at step k0 + k) and tests delta = max |omega_old - omega|
// against eps (:252) -- instead of ~8 small launches and a host sync per step (C1: 38 us of kernel in a
// 110 us step).  Update rules, rounded exactly as the host optimizer's separate tensor ops:
//   kind 0  Sga     omega += lr * grad          (optimizer.py:104-107)
//   kind 1  ExpSga  omega *= exp(lr * grad)     (optimizer.py:161-164; normalize = False)
// A warp stops at convergence (done = 1) or when the rates run out (the host relaunches with the next
// chunk); NaN ends the loop like the reference's `while delta > eps`.
// ---------------------------------------------------------------------------
struct IrlLoopArgs {
    double *theta;               // [B][S] in / out
    const double *lr;            // [B or 1][n_rates]
    size_t lr_stride;            // 0: one schedule for the whole batch
    int n_rates;
    int kind;
    double eps;
    int32_t *steps;              // [B] in / out: outer steps taken so far
    int32_t *done;               // [B] in / out: 1 = the loop has ended
    int32_t *last_counts;        // [B][2] or null: sweeps of the last gradient step
};

// ---------------------------------------------------------------------------
// Soft value iteration of ONE tiny world by FOUR warps (lane = state, warp = action), for the latency-bound
// case of a single causal problem (BASELINE configs[1]): one warp runs a sweep as ~300 dependent instructions
// (4 exp + 1 log per state, ncu: 54 % stall_wait, 1 240 cycles per sweep).  Here warp a computes only
// q_a = r + g P_a.v and exp(q_a - m); the four warps exchange q and the exponentials through shared memory (two
// bar.sync per sweep, buffers alternate by sweep parity) and every warp then forms the same v' = m + log(sum) --
// so each warp keeps its own full copy of v in registers, gathers by shuffle, and votes for itself.
// The expression tree per state is exactly succ_update<kOpSoftVI>: bitwise the one-warp kernel's results.
// ---------------------------------------------------------------------------
struct Cta4Smem {
    double q[2][4][32];
    double e[2][5][32];          // [..][4] = exp(phi - m)
    double pol[4][32];
};

__device__ __forceinline__ void cta4_soft_vi(const WarpWorld &wd, const StepBatch &bt, Cta4Smem &sm, const double r,
                                             const int lane, const int wa, double (&pol)[4], int &n_pol, int &st_pol) {
    constexpr int A = 4, K = 5;
    constexpr unsigned FULL = 0xffffffffu;
    const SuccArgs &s = bt.s;
    const bool act = wd.act;
    double x = kNegHuge;                                                         // maxent.py:323
    double q[A], qa = 0.0;
    double pra[K];                                                               // this warp's action row, in registers
#pragma unroll
    for (int j = 0; j < K; ++j) {
        pra[j] = wd.pr[0][j];
#pragma unroll
        for (int a = 1; a < A; ++a) pra[j] = (a == wa) ? wd.pr[a][j] : pra[j];
    }
    n_pol = 0;
    st_pol = IRLB200_ST_CONVERGED;
    const int limit = s.max_sweeps > 0 ? s.max_sweeps : 0x7fffffff;
    for (;;) {
        const int par = n_pol & 1;
        double dot = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) dot = fma(pra[j], __shfl_sync(FULL, x, wd.ix[j]), dot);
        qa = r + s.discount * dot;                                                // :329, this warp's action
        sm.q[par][wa][lane] = qa;
        __syncthreads();
#pragma unroll
        for (int a = 0; a < A; ++a) q[a] = sm.q[par][a][lane];
        double m = wd.c1;
#pragma unroll
        for (int a = 0; a < A; ++a) m = max_nan(m, q[a]);
        double xn;
        const bool finite = fabs(m) < INFINITY;
        // the branch may differ between lanes, never between the warps of one lane: every warp sees the same m
        if (finite) {
            sm.e[par][wa][lane] = exp(qa - m);
            if (wa == 0) sm.e[par][4][lane] = (wd.c1 != -INFINITY) ? exp(wd.c1 - m) : 0.0;
        }
        __syncthreads();
        if (finite) {
            double ssum = sm.e[par][4][lane];
#pragma unroll
            for (int a = 0; a < A; ++a) ssum += sm.e[par][a][lane];
            xn = m + log(ssum);
        } else {
            xn = wd.c1;                                                          // the reference's fold, verbatim
#pragma unroll 1
            for (int a = 0; a < A; ++a) xn = softmax2(xn, q[a]);
        }
        const double diff = fabs(xn - x);
        x = xn;
        ++n_pol;
        const bool nan = __any_sync(FULL, act && diff != diff);
        const bool gt = __any_sync(FULL, act && diff > s.eps);
        if (nan) { st_pol = IRLB200_ST_NONFINITE; break; }
        if (!gt) break;
        if (n_pol >= limit) { st_pol = IRLB200_ST_MAXSWEEPS; break; }
    }
    // policy of the last sweep (:341): q of that sweep against the new v; one action per warp, then shared
    sm.pol[wa][lane] = exp(qa - x);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < A; ++a) pol[a] = sm.pol[a][lane];
    __syncthreads();                                                             // sm.pol is rewritten next step
}

// forward pass of a warp's world from a given policy (the second half of warp_world_step)
__device__ __forceinline__ double warp_world_forward(const WarpWorld &wd, const StepBatch &bt, const double (&pol)[4],
                                                     int &n_svf, int &st_svf) {
    constexpr int A = 4, K = 5;
    constexpr unsigned FULL = 0xffffffffu;
    const SvfArgs &f = bt.f;
    const bool act = wd.act;
    double w[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < A; ++a) acc = fma(wd.pp[a][j], __shfl_sync(FULL, pol[a], wd.px[j]), acc);
        const int pterm = __shfl_sync(FULL, wd.is_term, wd.px[j]);
        w[j] = (pterm || !act) ? 0.0 : acc;
    }
    double d = 0.0;
    n_svf = 0;
    st_svf = IRLB200_ST_CONVERGED;
    const int limit = f.max_sweeps > 0 ? f.max_sweeps : 0x7fffffff;
    for (;;) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) acc = fma(w[j], __shfl_sync(FULL, d, wd.px[j]), acc);
        const double dn = wd.p0 + acc;
        const double diff = fabs(dn - d);
        d = dn;
        ++n_svf;
        const bool nan = __any_sync(FULL, act && diff != diff);
        const bool gt = __any_sync(FULL, act && diff > f.eps);
        if (nan) { st_svf = IRLB200_ST_NONFINITE; break; }
        if (!gt) break;
        if (n_svf >= limit) { st_svf = IRLB200_ST_MAXSWEEPS; break; }
    }
    return d;
}

// irl_causal's outer loop, one CTA of four warps per problem (see cta4_soft_vi); the forward pass and the
// optimizer step are computed by every warp redundantly (identical registers, no exchange), warp 0 writes.
__global__ void __launch_bounds__(128) irl_cta4_kernel(const StepBatch bt, const IrlLoopArgs lp, const int B) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ Cta4Smem sm;
    const int lane = threadIdx.x & 31, wa = threadIdx.x >> 5;
    const size_t b = blockIdx.x;
    if (b >= (size_t)B || lp.done[b]) return;                                    // uniform over the CTA
    const int S = bt.s.S;
    WarpWorld wd;
    warp_world_load<true>(wd, bt, b, lane);
    const int me = wd.act ? lane : 0;
    const double ef = wd.act ? bt.f.e_features[b * bt.ef_stride + me] : 0.0;
    const double *lr = lp.lr + b * lp.lr_stride;
    double theta = wd.act ? lp.theta[b * S + me] : 0.0;
    int steps = 0, done = 0, n_pol = 0, n_svf = 0;
    for (int k = 0; k < lp.n_rates; ++k) {
        double pol[4];
        int st_pol, st_svf;
        cta4_soft_vi(wd, bt, sm, theta, lane, wa, pol, n_pol, st_pol);
        const double d = warp_world_forward(wd, bt, pol, n_svf, st_svf);
        const double grad = ef - d;
        const double rate = __ldg(lr + k);
        const double old = theta;
        if (lp.kind == 0) theta = __dadd_rn(theta, __dmul_rn(rate, grad));
        else theta = __dmul_rn(theta, exp(__dmul_rn(rate, grad)));
        ++steps;
        const double diff = fabs(old - theta);
        const bool nan = __any_sync(FULL, wd.act && diff != diff);
        const bool gt = __any_sync(FULL, wd.act && diff > lp.eps);
        if (nan || !gt) { done = 1; break; }
    }
    if (wa == 0) {
        if (wd.act) lp.theta[b * S + me] = theta;
        if (lane == 0) {
            lp.steps[b] += steps;
            lp.done[b] = done;
            if (lp.last_counts) { lp.last_counts[2 * b] = n_pol; lp.last_counts[2 * b + 1] = n_svf; }
        }
    }
}

template <bool CAUSAL>
__global__ void __launch_bounds__(128) irl_warp_kernel(const StepBatch bt, const IrlLoopArgs lp, const int B) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t b = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= (size_t)B) return;
    if (lp.done[b]) return;
    const int S = bt.s.S;
    WarpWorld wd;
    warp_world_load<CAUSAL>(wd, bt, b, lane);
    const int me = wd.act ? lane : 0;
    const double ef = wd.act ? bt.f.e_features[b * bt.ef_stride + me] : 0.0;
    const double *lr = lp.lr + b * lp.lr_stride;
    double theta = wd.act ? lp.theta[b * S + me] : 0.0;
    int steps = 0, done = 0, n_pol = 0, n_svf = 0;
    for (int k = 0; k < lp.n_rates; ++k) {
        double pol[4];
        int st_pol, st_svf;
        const double d = warp_world_step<CAUSAL>(wd, bt, theta, pol, n_pol, st_pol, n_svf, st_svf);   // reward = omega (:244)
        const double grad = ef - d;                                                                   // :248
        const double rate = __ldg(lr + k);
        const double old = theta;
        if (lp.kind == 0) theta = __dadd_rn(theta, __dmul_rn(rate, grad));
        else theta = __dmul_rn(theta, exp(__dmul_rn(rate, grad)));
        ++steps;
        const double diff = fabs(old - theta);
        const bool nan = __any_sync(FULL, wd.act && diff != diff);
        const bool gt = __any_sync(FULL, wd.act && diff > lp.eps);
        if (nan || !gt) { done = 1; break; }                                                          // :240 / :252
    }
    if (wd.act) lp.theta[b * S + me] = theta;
    if (lane == 0) {
        lp.steps[b] += steps;
        lp.done[b] = done;
        if (lp.last_counts) { lp.last_counts[2 * b] = n_pol; lp.last_counts[2 * b + 1] = n_svf; }
    }
}

// ---------------------------------------------------------------------------
// Hand-tuned forward pass for the register-resident shape (A = 4, Kp = 5).
//
// Same arithmetic, same order as svf_phase<CtaTopo, 4, 5, SPT> -- results are
// bit-identical -- but the sweep body is stripped to what the FP64 pipe and the
// issue slots must do:
//   * the two iterate buffers sit STRIDE bytes apart (compile-time), the five
//     gather addresses of every owned state are precomputed 32-bit shared
//     addresses, so a sweep is 5 x (LDS.64 [addr + imm]; DFMA) per state with
//     no address arithmetic at all;
//   * states beyond S are padded with zero weights instead of predicated out;
//   * the stop rule is one DSETP per state accumulated in a predicate
//     (`!(|diff| <= eps)`, true for "greater" and for NaN) and one bar.red.or per
//     sweep; whether a surviving vote came from a non-finite iterate is checked
//     every 16 sweeps (such an iterate is sticky, so the loop ends within 16
//     sweeps of the reference's NaN exit; convergent runs stop on exactly the
//     reference's sweep).
// ---------------------------------------------------------------------------
// One sweep of the hand-tuned forward kernel: all gathers first, then SPT independent
// DFMA chains (the compiler interleaves them, hiding the 8-cycle DFMA latency), then the
// stores and the stop-rule predicates.  OFF_R / OFF_W select the iterate buffers.
template <int SPT, int OFF_R, int OFF_W>
__device__ __forceinline__ bool svf_fast_sweep(unsigned char *smem, const uint32_t (&ad)[SPT][5],
                                               const uint32_t (&own)[SPT], const double (&w)[SPT][5],
                                               const double (&p0r)[SPT], double (&cur)[SPT], double eps) {
    double v[SPT][5], x[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
        for (int j = 0; j < 5; ++j) v[k][j] = *reinterpret_cast<const double *>(smem + ad[k][j] + OFF_R);
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        double acc = fma(w[k][0], v[k][0], 0.0);
        acc = fma(w[k][1], v[k][1], acc);
        acc = fma(w[k][2], v[k][2], acc);
        acc = fma(w[k][3], v[k][3], acc);
        acc = fma(w[k][4], v[k][4], acc);
        x[k] = p0r[k] + acc;                                            // p_initial + sum   maxent.py:110
    }
    bool go = false;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        *reinterpret_cast<double *>(smem + own[k] + OFF_W) = x[k];
        go |= !(fabs(x[k] - cur[k]) <= eps);                            // |diff| > eps, or NaN
        cur[k] = x[k];
    }
    return go;
}

template <int SPT, int STRIDE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) svf_cta_fast_kernel(const SvfBatch bt) {
    constexpr int K = 5, A = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *flag = reinterpret_cast<int *>(smem_raw + 2 * STRIDE);

    SvfArgs a = bt.a;
    const size_t prob = svf_problem(bt);
    offset_svf(a, bt, prob);
    const int S = a.S, T = blockDim.x, tid = threadIdx.x;

    double w[SPT][K], p0r[SPT], cur[SPT];
    uint32_t ad[SPT][K], own[SPT];          // byte offsets into the first iterate buffer

#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int s = tid + k * T;
        const bool act = s < S;
        own[k] = 8u * (uint32_t)s;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int pred = act ? a.idx[(size_t)j * S + s] : s;
            double acc = 0.0;
            if (act) {
#pragma unroll
                for (int aa = 0; aa < A; ++aa)
                    acc = fma(__ldg(a.p + ((size_t)aa * K + j) * S + s), a.policy[(size_t)pred * A + aa], acc);
                if (a.term[pred]) acc = 0.0;
            }
            w[k][j] = acc;
            ad[k][j] = 8u * (uint32_t)pred;
        }
        p0r[k] = act ? a.p0[s] : 0.0;
        cur[k] = 0.0;
        *reinterpret_cast<double *>(smem_raw + own[k]) = 0.0;
        *reinterpret_cast<double *>(smem_raw + own[k] + STRIDE) = 0.0;
    }
    if (tid == 0) *flag = 0;
    __syncthreads();

    const double eps = a.eps;
    const int limit = a.max_sweeps > 0 ? a.max_sweeps : 0x7fffffff;
    int n = 0, status = IRLB200_ST_CONVERGED;
    for (;;) {
        const bool go = (n & 1) ? svf_fast_sweep<SPT, STRIDE, 0>(smem_raw, ad, own, w, p0r, cur, eps)
                                : svf_fast_sweep<SPT, 0, STRIDE>(smem_raw, ad, own, w, p0r, cur, eps);
        ++n;
        if (!__syncthreads_or(go ? 1 : 0)) break;                       // delta <= eps: converged
        if ((n & 15) == 0) {                                            // did a vote survive on NaN?
            bool bad = false;
#pragma unroll
            for (int k = 0; k < SPT; ++k) bad |= (cur[k] - cur[k]) != 0.0;   // NaN or +-inf iterate
            if (bad) *flag = 1;     // an infinite iterate makes the next diff inf - inf = NaN
            __syncthreads();
            if (*flag) { status = IRLB200_ST_NONFINITE; break; }
        }
        if (n >= limit) { status = IRLB200_ST_MAXSWEEPS; break; }
    }

#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int s = tid + k * T;
        if (s < S) {
            a.svf[s] = cur[k];
            if (a.grad) a.grad[s] = a.e_features[s] - cur[k];
        }
    }
    if (tid == 0) {
        if (bt.n_iter) bt.n_iter[prob * bt.out_stride] = n;
        if (bt.status) bt.status[prob * bt.out_stride] = status;
    }
}


}  // namespace irlb200
