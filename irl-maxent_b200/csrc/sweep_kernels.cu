// sweep_kernels.cu -- __global__ wrappers around phases.cuh and their launchers.
//
//   *_cta_kernel    grid = B problems, one CTA each (iterate in shared memory)
//   *_grid_kernel   one problem, cooperative persistent grid (iterate in L2/HBM)
//   step_cta_kernel fused gradient step: policy phase + forward phase, the
//                   policy stays in shared memory
#include <cooperative_groups.h>

#include <cstdio>
#include <mutex>

#include "host_util.h"
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
};

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
    offset_svf(a, bt, blockIdx.x);
    int *ni = bt.n_iter ? bt.n_iter + (size_t)blockIdx.x * bt.out_stride : nullptr;
    int *st = bt.status ? bt.status + (size_t)blockIdx.x * bt.out_stride : nullptr;
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
template <bool CAUSAL>
__global__ void __launch_bounds__(128) step_warp_kernel(const StepBatch bt, const int B) {
    constexpr int A = 4, K = 5;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t b = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= (size_t)B) return;                                   // whole warps leave: no block-level sync below
    SuccArgs s = bt.s;
    SvfArgs f = bt.f;
    const int S = s.S;
    const bool act = lane < S;
    const int me = act ? lane : 0;
    s.idx += b * bt.succ_idx_stride; s.p += b * bt.succ_p_stride; s.reward += b * S;
    if (s.phi) s.phi += b * bt.phi_stride;
    if (s.term) s.term += b * bt.term_stride;
    f.idx += b * bt.pred_idx_stride; f.p += b * bt.pred_p_stride; f.p0 += b * bt.p0_stride; f.term += b * bt.term_stride;

    // ---- policy pass -----------------------------------------------------------------------------
    int ix[K];
    double pr[A][K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        ix[j] = act ? s.idx[(size_t)j * S + me] : 0;
#pragma unroll
        for (int a = 0; a < A; ++a) pr[a][j] = act ? s.p[((size_t)a * K + j) * S + me] : 0.0;
    }
    const double r = act ? s.reward[me] : 0.0;
    const double c0 = CAUSAL ? r : exp(r);
    const double c1 = (CAUSAL && act) ? s.phi[me] : 0.0;
    double x = CAUSAL ? kNegHuge : ((act && s.term[me]) ? 1.0 : 0.0);
    double x_old = x;
    int n_pol = 0, st_pol = IRLB200_ST_CONVERGED;
    auto gather_update = [&](double xin, double *q) {
        double xv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) xv[j] = __shfl_sync(FULL, xin, ix[j]);
        return succ_update<CAUSAL ? kOpSoftVI : kOpBackward, 4>(
            A, K, [&](int a, int j) { return pr[a][j]; }, [&](int j) { return xv[j]; }, c0, c1, s.discount, 0, q);
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
        double mr = fabs(r);
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
    double pol[A];
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
    if (act && bt.policy_out) {
#pragma unroll
        for (int a = 0; a < A; ++a) bt.policy_out[(b * S + me) * A + a] = pol[a];
    }

    // ---- forward pass ------------------------------------------------------------------------------
    int px[K];
    double w[K];
    const int is_term = (act && f.term[me]) ? 1 : 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        px[j] = act ? f.idx[(size_t)j * S + me] : 0;
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const double pa = __shfl_sync(FULL, pol[a], px[j]);                 // policy[pred_j, a]
            const double pp = act ? __ldg(f.p + ((size_t)a * K + j) * S + me) : 0.0;
            acc = fma(pp, pa, acc);
        }
        const int pterm = __shfl_sync(FULL, is_term, px[j]);
        w[j] = (pterm || !act) ? 0.0 : acc;
    }
    const double p0 = act ? f.p0[me] : 0.0;
    double d = 0.0;
    int n_svf = 0, st_svf = IRLB200_ST_CONVERGED;
    const int limit = f.max_sweeps > 0 ? f.max_sweeps : 0x7fffffff;
    for (;;) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) acc = fma(w[j], __shfl_sync(FULL, d, px[j]), acc);
        const double dn = p0 + acc;
        const double diff = fabs(dn - d);
        d = dn;
        ++n_svf;
        const bool nan = __any_sync(FULL, act && diff != diff);
        const bool gt = __any_sync(FULL, act && diff > f.eps);
        if (nan) { st_svf = IRLB200_ST_NONFINITE; break; }
        if (!gt) break;
        if (n_svf >= limit) { st_svf = IRLB200_ST_MAXSWEEPS; break; }
    }
    if (act) {
        f.svf[b * S + me] = d;
        if (f.grad) f.grad[b * S + me] = f.e_features[b * bt.ef_stride + me] - d;
    }
    if (lane == 0) {
        if (bt.n_iter) { bt.n_iter[2 * b] = n_pol; bt.n_iter[2 * b + 1] = n_svf; }
        if (bt.status) { bt.status[2 * b] = st_pol; bt.status[2 * b + 1] = st_svf; }
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
    offset_svf(a, bt, blockIdx.x);
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
        if (bt.n_iter) bt.n_iter[(size_t)blockIdx.x * bt.out_stride] = n;
        if (bt.status) bt.status[(size_t)blockIdx.x * bt.out_stride] = status;
    }
}

// ---------------------------------------------------------------------------
// Stencil-tiled forward pass for grid worlds (predecessor offsets within {-n,-1,0,+1,+n}).
//
// The generic gather above is bound by the shared-memory datapath: 5 LDS.64 per state =
// 384 wavefronts per 1 024-state sweep against 141 FP64-pipe cycles.  Here a thread owns a
// TY x TX tile of grid cells and keeps their iterate values in registers; only the halo of
// the tile (TX cells above / below, TY cells left / right) is read from shared memory:
// (2 TX + 2 TY) / (TX TY) loads per state (1.5 for 2 x 4, 1.0 for 4 x 4).  Every thread
// publishes its tile in a private slot of PITCH = TX TY + 1 doubles; the odd pitch makes
// the 64-bit accesses of a half-warp hit distinct banks, and a cell's offset inside the
// slot is an immediate.
//
// Arithmetic: per state the same ascending-neighbour FMA chain as the ELL kernels
// (s-n, s-1, s, s+1, s+n); a neighbour that is absent from the table enters as
// fma(0, v, acc) == acc, so results are bit-identical to svf_cta_fast_kernel.
// ---------------------------------------------------------------------------
template <int TY, int TX, int MAXT>
struct Grid5Cfg {
    static constexpr int C = TY * TX;
    static constexpr int PITCH = C + 1 + ((C + 1) % 2 == 0 ? 1 : 0);     // odd number of doubles
    static constexpr int STRIDE = MAXT * PITCH * 8;                      // bytes between the two buffers
};

template <int TY, int TX, int MAXT, int OFF_R, int OFF_W>
__device__ __forceinline__ void svf_grid5_sweep(unsigned char *smem, uint32_t own, uint32_t nb_up, uint32_t nb_dn,
                                                uint32_t nb_lf, uint32_t nb_rt, const double (&w)[TY * TX][5],
                                                const double (&p0r)[TY * TX], const double (&cur)[TY * TX],
                                                double (&x)[TY * TX]) {
    double up[TX], dn[TX], lf[TY], rt[TY];
#pragma unroll
    for (int ix = 0; ix < TX; ++ix) {
        up[ix] = *reinterpret_cast<const double *>(smem + nb_up + 8 * ((TY - 1) * TX + ix) + OFF_R);
        dn[ix] = *reinterpret_cast<const double *>(smem + nb_dn + 8 * ix + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy) {
        lf[iy] = *reinterpret_cast<const double *>(smem + nb_lf + 8 * (iy * TX + TX - 1) + OFF_R);
        rt[iy] = *reinterpret_cast<const double *>(smem + nb_rt + 8 * (iy * TX) + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const double v_up = iy > 0 ? cur[c - TX] : up[ix];
            const double v_lf = ix > 0 ? cur[c - 1] : lf[iy];
            const double v_rt = ix < TX - 1 ? cur[c + 1] : rt[iy];
            const double v_dn = iy < TY - 1 ? cur[c + TX] : dn[ix];
            double acc = fma(w[c][0], v_up, 0.0);
            acc = fma(w[c][1], v_lf, acc);
            acc = fma(w[c][2], cur[c], acc);
            acc = fma(w[c][3], v_rt, acc);
            acc = fma(w[c][4], v_dn, acc);
            x[c] = p0r[c] + acc;                                        // p_initial + sum   maxent.py:110
        }
#pragma unroll
    for (int c = 0; c < TY * TX; ++c) *reinterpret_cast<double *>(smem + own + 8 * c + OFF_W) = x[c];
}

template <int TY, int TX, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) svf_grid5_kernel(const SvfBatch bt, const int n) {
    using Cfg = Grid5Cfg<TY, TX, MAXT>;
    constexpr int C = Cfg::C, K = 5, A = 4, STRIDE = Cfg::STRIDE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *flag = reinterpret_cast<int *>(smem_raw + 2 * STRIDE);

    SvfArgs a = bt.a;
    offset_svf(a, bt, blockIdx.x);
    const int S = a.S, tid = threadIdx.x;
    const int ntx = n / TX, nty = n / TY;
    const bool live = tid < ntx * nty;
    const int tx = live ? tid % ntx : 0, ty = live ? tid / ntx : 0;

    double w[C][5], p0r[C], cur[C];
    const uint32_t slot = 8u * Cfg::PITCH;
    const uint32_t own = slot * tid;
    const uint32_t nb_up = (live && ty > 0) ? own - slot * ntx : own;
    const uint32_t nb_dn = (live && ty < nty - 1) ? own + slot * ntx : own;
    const uint32_t nb_lf = (live && tx > 0) ? own - slot : own;
    const uint32_t nb_rt = (live && tx < ntx - 1) ? own + slot : own;

#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const int s = (ty * TY + iy) * n + tx * TX + ix;
#pragma unroll
            for (int k = 0; k < 5; ++k) w[c][k] = 0.0;
            if (live) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int pred = a.idx[(size_t)j * S + s];
                    double acc = 0.0;
#pragma unroll
                    for (int aa = 0; aa < A; ++aa)
                        acc = fma(__ldg(a.p + ((size_t)aa * K + j) * S + s), a.policy[(size_t)pred * A + aa], acc);
                    if (a.term[pred]) acc = 0.0;
                    const int off = pred - s;
                    // at most one table entry per offset carries weight; padding entries add +0
                    w[c][0] += (off == -n) ? acc : 0.0;
                    w[c][1] += (off == -1) ? acc : 0.0;
                    w[c][2] += (off == 0) ? acc : 0.0;
                    w[c][3] += (off == 1) ? acc : 0.0;
                    w[c][4] += (off == n) ? acc : 0.0;
                }
            }
            p0r[c] = live ? a.p0[s] : 0.0;
            cur[c] = 0.0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c) = 0.0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c + STRIDE) = 0.0;
        }
    if (tid == 0) *flag = 0;
    __syncthreads();

    // Stop rule (`while delta > eps`, maxent.py:108-112), exact but cheap: a sweep must continue as
    // soon as ANY state moved by more than eps, so each thread first votes with one cell of its tile
    // only; the full per-cell test (and a second bar.red) runs just in the sweeps where that sampled
    // vote finds nothing -- the last few hundred of ~10^4..10^5.  Stopping always needs the full test.
    const double eps = a.eps;
    const int limit = a.max_sweeps > 0 ? a.max_sweeps : 0x7fffffff;
    int nsw = 0, status = IRLB200_ST_CONVERGED;
    for (;;) {
        double x[C];
        if (nsw & 1) svf_grid5_sweep<TY, TX, MAXT, STRIDE, 0>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, p0r, cur, x);
        else svf_grid5_sweep<TY, TX, MAXT, 0, STRIDE>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, p0r, cur, x);
        ++nsw;
        bool stop = false;
        if (!__syncthreads_or(!(fabs(x[0] - cur[0]) <= eps) ? 1 : 0)) {
            bool go = false;
#pragma unroll
            for (int c = 1; c < C; ++c) go |= !(fabs(x[c] - cur[c]) <= eps);   // |diff| > eps, or NaN
            stop = !__syncthreads_or(go ? 1 : 0);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) cur[c] = x[c];
        if (stop) break;                                                // delta <= eps: converged
        if ((nsw & 15) == 0) {
            bool bad = false;
#pragma unroll
            for (int c = 0; c < C; ++c) bad |= (cur[c] - cur[c]) != 0.0;
            if (bad) *flag = 1;
            __syncthreads();
            if (*flag) { status = IRLB200_ST_NONFINITE; break; }
        }
        if (nsw >= limit) { status = IRLB200_ST_MAXSWEEPS; break; }
    }

    if (live) {
#pragma unroll
        for (int iy = 0; iy < TY; ++iy)
#pragma unroll
            for (int ix = 0; ix < TX; ++ix) {
                const int c = iy * TX + ix;
                const int s = (ty * TY + iy) * n + tx * TX + ix;
                a.svf[s] = cur[c];
                if (a.grad) a.grad[s] = a.e_features[s] - cur[c];
            }
    }
    if (tid == 0) {
        if (bt.n_iter) bt.n_iter[(size_t)blockIdx.x * bt.out_stride] = nsw;
        if (bt.status) bt.status[(size_t)blockIdx.x * bt.out_stride] = status;
    }
}

// ---------------------------------------------------------------------------
// Cluster variant of the stencil-tiled forward pass: ONE world spread over a thread-block
// cluster (up to 16 CTAs), for worlds too large for one CTA but small enough that a grid
// barrier (~1.5 us) would dominate the sweep (BASELINE configs[2], 128 x 128).
// CTA c of the cluster owns the tile rows [c R, (c+1) R); the up / down halo of its first /
// last tile row is read straight from the neighbouring CTA's shared memory (DSMEM), the
// sweep fence is barrier.cluster (arrive.release / wait.acquire), and the stop rule is
// all-reduced by stamping the sweep number into every CTA's vote word with remote
// shared-memory stores before the cluster barrier (double-buffered by sweep parity).
// Same tile arithmetic as svf_grid5_kernel: bitwise identical results.
// ---------------------------------------------------------------------------
namespace cgx = cooperative_groups;

template <int TY, int TX, int OFF_R, int OFF_W>
__device__ __forceinline__ void svf_grid5_cluster_sweep(unsigned char *smem, uint32_t own, const unsigned char *up_p,
                                                        const unsigned char *dn_p, uint32_t nb_lf, uint32_t nb_rt,
                                                        const double (&w)[TY * TX][5], const double (&p0r)[TY * TX],
                                                        const double (&cur)[TY * TX], double (&x)[TY * TX]) {
    double up[TX], dn[TX], lf[TY], rt[TY];
#pragma unroll
    for (int ix = 0; ix < TX; ++ix) {
        up[ix] = *reinterpret_cast<const double *>(up_p + 8 * ((TY - 1) * TX + ix) + OFF_R);
        dn[ix] = *reinterpret_cast<const double *>(dn_p + 8 * ix + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy) {
        lf[iy] = *reinterpret_cast<const double *>(smem + nb_lf + 8 * (iy * TX + TX - 1) + OFF_R);
        rt[iy] = *reinterpret_cast<const double *>(smem + nb_rt + 8 * (iy * TX) + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const double v_up = iy > 0 ? cur[c - TX] : up[ix];
            const double v_lf = ix > 0 ? cur[c - 1] : lf[iy];
            const double v_rt = ix < TX - 1 ? cur[c + 1] : rt[iy];
            const double v_dn = iy < TY - 1 ? cur[c + TX] : dn[ix];
            double acc = fma(w[c][0], v_up, 0.0);
            acc = fma(w[c][1], v_lf, acc);
            acc = fma(w[c][2], cur[c], acc);
            acc = fma(w[c][3], v_rt, acc);
            acc = fma(w[c][4], v_dn, acc);
            x[c] = p0r[c] + acc;                                        // p_initial + sum   maxent.py:110
        }
#pragma unroll
    for (int c = 0; c < TY * TX; ++c) *reinterpret_cast<double *>(smem + own + 8 * c + OFF_W) = x[c];
}

// cluster-wide OR of a per-thread predicate: stamp, barrier.cluster, compare.  `word` points at this
// CTA's vote words [2]; every CTA's copy is written by every voting warp (one lane per target CTA).
__device__ __forceinline__ bool cluster_any(cgx::cluster_group &cl, int *word, int stamp, bool pred, int ncta) {
    int *slot = word + (stamp & 1);
    const unsigned any = __ballot_sync(0xffffffffu, pred);
    const int lane = threadIdx.x & 31;
    if (any && lane < ncta) *cl.map_shared_rank(slot, lane) = stamp;
    cl.sync();
    return *slot == stamp;
}

template <int TY, int TX, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) svf_grid5_cluster_kernel(const SvfBatch bt, const int n, const int R) {
    using Cfg = Grid5Cfg<TY, TX, MAXT>;
    constexpr int C = Cfg::C, K = 5, A = 4, STRIDE = Cfg::STRIDE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *votes = reinterpret_cast<int *>(smem_raw + 2 * STRIDE);      // [0..1] continue stamps, [2..3] non-finite stamps
    cgx::cluster_group cl = cgx::this_cluster();
    const int ncta = (int)cl.num_blocks(), crank = (int)cl.block_rank();

    SvfArgs a = bt.a;
    offset_svf(a, bt, blockIdx.x / ncta);
    const int S = a.S, tid = threadIdx.x;
    const int ntx = n / TX;
    const bool live = tid < ntx * R;
    const int tx = live ? tid % ntx : 0, lty = live ? tid / ntx : 0;
    const int gty = crank * R + lty, nty = n / TY;

    double w[C][5], p0r[C], cur[C];
    const uint32_t slot = 8u * Cfg::PITCH;
    const uint32_t own = slot * tid;
    const uint32_t nb_lf = (live && tx > 0) ? own - slot : own;
    const uint32_t nb_rt = (live && tx < ntx - 1) ? own + slot : own;
    const unsigned char *up_p = smem_raw + own, *dn_p = smem_raw + own;
    if (live && lty > 0) up_p = smem_raw + own - slot * ntx;
    else if (live && gty > 0) up_p = cl.map_shared_rank(smem_raw, crank - 1) + slot * ((R - 1) * ntx + tx);
    if (live && lty < R - 1) dn_p = smem_raw + own + slot * ntx;
    else if (live && gty < nty - 1) dn_p = cl.map_shared_rank(smem_raw, crank + 1) + slot * tx;

#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const int s = (gty * TY + iy) * n + tx * TX + ix;
#pragma unroll
            for (int k = 0; k < 5; ++k) w[c][k] = 0.0;
            if (live) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int pred = a.idx[(size_t)j * S + s];
                    double acc = 0.0;
#pragma unroll
                    for (int aa = 0; aa < A; ++aa)
                        acc = fma(__ldg(a.p + ((size_t)aa * K + j) * S + s), a.policy[(size_t)pred * A + aa], acc);
                    if (a.term[pred]) acc = 0.0;
                    const int off = pred - s;
                    w[c][0] += (off == -n) ? acc : 0.0;
                    w[c][1] += (off == -1) ? acc : 0.0;
                    w[c][2] += (off == 0) ? acc : 0.0;
                    w[c][3] += (off == 1) ? acc : 0.0;
                    w[c][4] += (off == n) ? acc : 0.0;
                }
            }
            p0r[c] = live ? a.p0[s] : 0.0;
            cur[c] = 0.0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c) = 0.0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c + STRIDE) = 0.0;
        }
    if (tid < 4) votes[tid] = 0;
    cl.sync();

    const double eps = a.eps;
    const int limit = a.max_sweeps > 0 ? a.max_sweeps : 0x7fffffff;
    int nsw = 0, status = IRLB200_ST_CONVERGED;
    for (;;) {
        double x[C];
        if (nsw & 1) svf_grid5_cluster_sweep<TY, TX, STRIDE, 0>(smem_raw, own, up_p, dn_p, nb_lf, nb_rt, w, p0r, cur, x);
        else svf_grid5_cluster_sweep<TY, TX, 0, STRIDE>(smem_raw, own, up_p, dn_p, nb_lf, nb_rt, w, p0r, cur, x);
        ++nsw;
        bool stop = false;
        // sampled vote first (one cell per tile), full test only when it finds nothing -- see svf_grid5_kernel
        if (!cluster_any(cl, votes, 2 * nsw, !(fabs(x[0] - cur[0]) <= eps), ncta)) {
            bool go = false;
#pragma unroll
            for (int c = 1; c < C; ++c) go |= !(fabs(x[c] - cur[c]) <= eps);
            stop = !cluster_any(cl, votes, 2 * nsw + 1, go, ncta);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) cur[c] = x[c];
        if (stop) break;
        if ((nsw & 15) == 0) {
            bool bad = false;
#pragma unroll
            for (int c = 0; c < C; ++c) bad |= (cur[c] - cur[c]) != 0.0;
            if (cluster_any(cl, votes + 2, nsw, bad, ncta)) { status = IRLB200_ST_NONFINITE; break; }
        }
        if (nsw >= limit) { status = IRLB200_ST_MAXSWEEPS; break; }
    }

    if (live) {
#pragma unroll
        for (int iy = 0; iy < TY; ++iy)
#pragma unroll
            for (int ix = 0; ix < TX; ++ix) {
                const int c = iy * TX + ix;
                const int s = (gty * TY + iy) * n + tx * TX + ix;
                a.svf[s] = cur[c];
                if (a.grad) a.grad[s] = a.e_features[s] - cur[c];
            }
    }
    if (tid == 0 && crank == 0) {
        const size_t wb = blockIdx.x / ncta;
        if (bt.n_iter) bt.n_iter[wb * bt.out_stride] = nsw;
        if (bt.status) bt.status[wb * bt.out_stride] = status;
    }
    cl.sync();      // no CTA may exit while a neighbour can still read its shared memory
}

// ---------------------------------------------------------------------------
// Push variant of the cluster forward pass (the default): no barrier.cluster in the loop.
//
// barrier.cluster (~430 cycles) followed by a dependent DSMEM load cost ~1 800 cycles per sweep at
// 128 x 128 (0.95 us); the arithmetic of a sweep is ~200-350.  Here every transfer is a one-way
// st.async -- a remote shared-memory store that decrements the transaction count of an mbarrier in
// the DESTINATION CTA (one flight, ~320 cycles measured, scripts/ubench_cluster.cu):
//   * halo rows: after a sweep, the threads of a CTA's first / last tile row write their boundary
//     cells straight into the neighbouring CTA's halo buffer; the sweep itself reads local shared
//     memory only.
//   * stop rule: every WARP votes on its own (sampled cell first, full test when the sample finds
//     nothing -- no CTA-wide reduction in front of the flight) and sends its vote word to the vote
//     table of every CTA of the cluster.
// Each CTA then waits on ONE mbarrier per sweep for (halo bytes + 4 ncta nwarps vote bytes), ORs the
// vote table and decides -- every CTA sees the same table, so all take the same decision on the same
// sweep, exactly where `while delta > eps` (maxent.py:108-112) stops.  A plain bar.sync orders the
// local tile slots; it overlaps the flight.
// (Ordinary remote stores instead of st.async are tracked by the sender's next release / bar.sync,
// which then waits for the store's round trip, ~650 cycles; a lagged all-reduce with a one-sweep
// rollback was measured too and lost: two mbarrier waits per sweep cost more than the flight saved.)
//
// Hazards: buffers and barriers are double-buffered by sweep parity.  A peer can only write parity q
// of sweep j+2 after it has seen the votes of ALL warps of this CTA for sweep j+1, each sent after
// that warp's last read of parity q.  Transaction bytes that land before the local
// arrive.expect_tx are legal (the phase cannot complete before the one expected arrival).
// Same tile arithmetic as svf_grid5_kernel: bitwise identical results, identical counts.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// false: the barrier did not complete within ~4 s (a peer CTA is gone)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 8000000000ll) return false;
    return true;
}
__device__ __forceinline__ void st_async_2f64(uint32_t raddr, double v0, double v1, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "d"(v0), "d"(v1), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_u32(uint32_t raddr, uint32_t v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];"
                 ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}

template <int TY, int TX, int MAXT>
struct PushCfg {
    using G = Grid5Cfg<TY, TX, MAXT>;
    static constexpr int kMaxN = 128;                                     // widest grid row
    static constexpr int kMaxCta = 16;
    static constexpr int NWARP = MAXT / 32;
    static_assert(kMaxCta * NWARP <= 128, "vote table: one uint4 per lane");
    static constexpr int SLOTS = MAXT * G::PITCH * 8;                     // tile slots of one parity
    static constexpr int HALO = 2 * kMaxN * 8;                            // [up row | down row]
    static constexpr int VOTES = 512;                                     // uint32[cta][warp]: bad << 1 | go
    static constexpr int STRIDE = SLOTS + HALO + VOTES;                   // bytes between the parities
    static constexpr int MBAR = 2 * STRIDE;                               // uint64[2], by parity
    static constexpr int BYTES = MBAR + 16;
};

// One sweep of iteration j (P = j & 1, compile time so that every shared-memory offset is an
// immediate): decide on sweep j-1, then sweep, push, vote.  Returns kContinue or a final status.
template <int TY, int TX, int MAXT, int P>
__device__ __forceinline__ int svf_push_iter(unsigned char *smem, const uint32_t sbase, const uint32_t own,
                                             const uint32_t nb_up, const uint32_t nb_dn, const uint32_t nb_lf,
                                             const uint32_t nb_rt, const uint32_t push_up, const uint32_t push_dn,
                                             const uint32_t bar_up, const uint32_t bar_dn, const uint32_t vote_dst,
                                             const uint32_t vote_bar, const uint32_t expect, const int ncta,
                                             const double (&w)[TY * TX][5], const double (&p0r)[TY * TX],
                                             double (&cur)[TY * TX], const double eps, const int limit, int &nsw) {
    using Cfg = PushCfg<TY, TX, MAXT>;
    constexpr int C = TY * TX, OFF_R = P * Cfg::STRIDE, OFF_W = (P ^ 1) * Cfg::STRIDE;
    const int j = nsw;                                                  // 0-based index of this sweep
    // arm the barrier that collects the rows and votes of THIS sweep (its previous phase, sweep j-2,
    // completed before this thread left iteration j-1's wait)
    if (threadIdx.x == 0) mbar_arrive_expect_tx(sbase + Cfg::MBAR + 8 * (P ^ 1), expect);
    if (j > 0) {
        // rows and votes of sweep j-1
        if (!mbar_wait(sbase + Cfg::MBAR + 8 * P, ((j - 1) >> 1) & 1)) return IRLB200_ST_ABORTED;
        const uint4 v = *reinterpret_cast<const uint4 *>(smem + Cfg::SLOTS + Cfg::HALO + 16 * (threadIdx.x & 31) + OFF_R);
        const unsigned all = __reduce_or_sync(0xffffffffu, v.x | v.y | v.z | v.w);
        if (!(all & 1u)) return IRLB200_ST_CONVERGED;                   // delta <= eps everywhere
        if (all & 2u) return IRLB200_ST_NONFINITE;
        if (j >= limit) return IRLB200_ST_MAXSWEEPS;
    }
    double up[TX], dn[TX], lf[TY], rt[TY], x[C];
#pragma unroll
    for (int ix = 0; ix < TX; ++ix) {
        up[ix] = *reinterpret_cast<const double *>(smem + nb_up + 8 * ((TY - 1) * TX + ix) + OFF_R);
        dn[ix] = *reinterpret_cast<const double *>(smem + nb_dn + 8 * ix + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy) {
        lf[iy] = *reinterpret_cast<const double *>(smem + nb_lf + 8 * (iy * TX + TX - 1) + OFF_R);
        rt[iy] = *reinterpret_cast<const double *>(smem + nb_rt + 8 * (iy * TX) + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const double v_up = iy > 0 ? cur[c - TX] : up[ix];
            const double v_lf = ix > 0 ? cur[c - 1] : lf[iy];
            const double v_rt = ix < TX - 1 ? cur[c + 1] : rt[iy];
            const double v_dn = iy < TY - 1 ? cur[c + TX] : dn[ix];
            double acc = fma(w[c][0], v_up, 0.0);
            acc = fma(w[c][1], v_lf, acc);
            acc = fma(w[c][2], cur[c], acc);
            acc = fma(w[c][3], v_rt, acc);
            acc = fma(w[c][4], v_dn, acc);
            x[c] = p0r[c] + acc;                                        // p_initial + sum   maxent.py:110
        }
    // boundary rows first: they have the longest way to go
    if (push_up) {
#pragma unroll
        for (int ix = 0; ix < TX; ix += 2) st_async_2f64(push_up + 8 * ix + OFF_W, x[ix], x[ix + 1], bar_up + 8 * (P ^ 1));
    }
    if (push_dn) {
#pragma unroll
        for (int ix = 0; ix < TX; ix += 2)
            st_async_2f64(push_dn + 8 * ix + OFF_W, x[(TY - 1) * TX + ix], x[(TY - 1) * TX + ix + 1], bar_dn + 8 * (P ^ 1));
    }
    nsw = j + 1;
    // this warp's vote: sampled cell first, full test only when the sample finds nothing
    unsigned vote = 1;
    if (!__any_sync(0xffffffffu, !(fabs(x[0] - cur[0]) <= eps))) {
        bool go = false;
#pragma unroll
        for (int c = 1; c < C; ++c) go |= !(fabs(x[c] - cur[c]) <= eps);   // |diff| > eps, or NaN
        vote = __any_sync(0xffffffffu, go) ? 1u : 0u;
    }
    if ((nsw & 15) == 0) {
        bool bad = false;
#pragma unroll
        for (int c = 0; c < C; ++c) bad |= (x[c] - x[c]) != 0.0;
        if (__any_sync(0xffffffffu, bad)) vote |= 2u;
    }
    if ((threadIdx.x & 31) < ncta) st_async_u32(vote_dst + OFF_W, vote, vote_bar + 8 * (P ^ 1));
#pragma unroll
    for (int c = 0; c < C; ++c) {
        *reinterpret_cast<double *>(smem + own + 8 * c + OFF_W) = x[c];
        cur[c] = x[c];
    }
    __syncthreads();                                                    // local tile slots of sweep j
    return kContinue;
}

template <int TY, int TX, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) svf_grid5_push_kernel(const SvfBatch bt, const int n, const int R) {
    using Cfg = PushCfg<TY, TX, MAXT>;
    constexpr int C = TY * TX, K = 5, A = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cgx::cluster_group cl = cgx::this_cluster();
    const int ncta = (int)cl.num_blocks(), crank = (int)cl.block_rank();
    const uint32_t sbase = smem_u32(smem_raw);

    SvfArgs a = bt.a;
    offset_svf(a, bt, blockIdx.x / ncta);
    const int S = a.S, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int ntx = n / TX;
    const bool live = tid < ntx * R;
    const int tx = live ? tid % ntx : 0, lty = live ? tid / ntx : 0;
    const int gty = crank * R + lty, nty = n / TY;

    double w[C][5], p0r[C], cur[C];
    const uint32_t slot = 8u * Cfg::G::PITCH;
    const uint32_t own = slot * tid;
    const uint32_t nb_lf = (live && tx > 0) ? own - slot : own;
    const uint32_t nb_rt = (live && tx < ntx - 1) ? own + slot : own;
    // up / down neighbours: another tile of this CTA, or the halo row received from the next CTA
    uint32_t nb_up = own, nb_dn = own, push_up = 0, push_dn = 0, bar_up = 0, bar_dn = 0;
    if (live && lty > 0) nb_up = own - slot * ntx;
    else if (live && gty > 0) {
        nb_up = Cfg::SLOTS + 8u * (tx * TX) - 8u * ((TY - 1) * TX);                 // halo "up" row
        push_up = mapa_u32(sbase + Cfg::SLOTS + 8u * (Cfg::kMaxN + tx * TX), crank - 1);   // its "down" row
        bar_up = mapa_u32(sbase + Cfg::MBAR, crank - 1);
    }
    if (live && lty < R - 1) nb_dn = own + slot * ntx;
    else if (live && gty < nty - 1) {
        nb_dn = Cfg::SLOTS + 8u * (Cfg::kMaxN + tx * TX);
        push_dn = mapa_u32(sbase + Cfg::SLOTS + 8u * (tx * TX), crank + 1);
        bar_dn = mapa_u32(sbase + Cfg::MBAR, crank + 1);
    }
    // votes: lane l < ncta of every warp writes the warp's word into CTA l's table
    const uint32_t vote_dst = lane < ncta ? mapa_u32(sbase + Cfg::SLOTS + Cfg::HALO + 4u * (crank * nwarp + warp), lane) : 0;
    const uint32_t vote_bar = lane < ncta ? mapa_u32(sbase + Cfg::MBAR, lane) : 0;
    const uint32_t expect = 4u * ncta * nwarp + (crank > 0 ? 8u * n : 0u) + (crank < ncta - 1 ? 8u * n : 0u);

#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const int s = (gty * TY + iy) * n + tx * TX + ix;
#pragma unroll
            for (int k = 0; k < 5; ++k) w[c][k] = 0.0;
            if (live) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int pred = a.idx[(size_t)j * S + s];
                    double acc = 0.0;
#pragma unroll
                    for (int aa = 0; aa < A; ++aa)
                        acc = fma(__ldg(a.p + ((size_t)aa * K + j) * S + s), a.policy[(size_t)pred * A + aa], acc);
                    if (a.term[pred]) acc = 0.0;
                    const int off = pred - s;
                    w[c][0] += (off == -n) ? acc : 0.0;
                    w[c][1] += (off == -1) ? acc : 0.0;
                    w[c][2] += (off == 0) ? acc : 0.0;
                    w[c][3] += (off == 1) ? acc : 0.0;
                    w[c][4] += (off == n) ? acc : 0.0;
                }
            }
            p0r[c] = live ? a.p0[s] : 0.0;
            cur[c] = 0.0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c) = 0.0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c + Cfg::STRIDE) = 0.0;
        }
    for (int i = tid; i < (Cfg::HALO + Cfg::VOTES) / 8; i += blockDim.x) {
        *reinterpret_cast<double *>(smem_raw + Cfg::SLOTS + 8 * i) = 0.0;
        *reinterpret_cast<double *>(smem_raw + Cfg::SLOTS + 8 * i + Cfg::STRIDE) = 0.0;
    }
    if (tid == 0) {
        mbar_init(sbase + Cfg::MBAR, 1);
        mbar_init(sbase + Cfg::MBAR + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cl.sync();

    const double eps = a.eps;
    const int limit = a.max_sweeps > 0 ? a.max_sweeps : 0x7fffffff;
    int nsw = 0, status;
    for (;;) {
        status = svf_push_iter<TY, TX, MAXT, 0>(smem_raw, sbase, own, nb_up, nb_dn, nb_lf, nb_rt, push_up, push_dn, bar_up,
                                                bar_dn, vote_dst, vote_bar, expect, ncta, w, p0r, cur, eps, limit, nsw);
        if (status != kContinue) break;
        status = svf_push_iter<TY, TX, MAXT, 1>(smem_raw, sbase, own, nb_up, nb_dn, nb_lf, nb_rt, push_up, push_dn, bar_up,
                                                bar_dn, vote_dst, vote_bar, expect, ncta, w, p0r, cur, eps, limit, nsw);
        if (status != kContinue) break;
    }
    // Every decision is taken after ALL bytes of the sweep it judges have landed in every CTA, so nothing
    // is in flight at this point -- except the barrier tid 0 armed for the sweep that was not run.

    if (live) {
#pragma unroll
        for (int iy = 0; iy < TY; ++iy)
#pragma unroll
            for (int ix = 0; ix < TX; ++ix) {
                const int c = iy * TX + ix;
                const int s = (gty * TY + iy) * n + tx * TX + ix;
                a.svf[s] = cur[c];
                if (a.grad) a.grad[s] = a.e_features[s] - cur[c];
            }
    }
    if (tid == 0 && crank == 0) {
        const size_t wb = blockIdx.x / ncta;
        if (bt.n_iter) bt.n_iter[wb * bt.out_stride] = nsw;
        if (bt.status) bt.status[wb * bt.out_stride] = status;
    }
    cl.sync();      // keep every CTA's shared memory alive until all peers are done with it
}

// ---------------------------------------------------------------------------
// Stencil-tiled non-causal backward pass (local_action_probabilities, maxent.py:119-159).
//
// All but the last of the n_sweeps partition sweeps only carry zs forward, and
//     zs'[s] = sum_a er[s] * sum_j P[s,j,a] * zs[j]  =  sum_j (er[s] * sum_a P[s,j,a]) * zs[j]
// is a 5-weight stencil exactly like the forward sweep: 5 FMA per state instead of 20 FMA +
// 4 MUL + 3 ADD (the FP64 pipe is what bounds these kernels).  The merged weights are formed
// once, so the iterate differs from the reference's by rounding only (~1e-15 relative, checked
// against the reference fixtures to 1e-10).  The LAST sweep is evaluated exactly as the
// reference does (per-action za = er * P_a.dot(zs), zs = za.sum, policy = za / zs, :155-159)
// from the ELL rows.  Range extension: exact power-of-two rescale every R sweeps.
// ---------------------------------------------------------------------------
template <int TY, int TX, int MAXT, int OFF_R, int OFF_W>
__device__ __forceinline__ double lin_grid5_sweep(unsigned char *smem, uint32_t own, uint32_t nb_up, uint32_t nb_dn,
                                                  uint32_t nb_lf, uint32_t nb_rt, const double (&w)[TY * TX][5],
                                                  double (&cur)[TY * TX]) {
    double up[TX], dn[TX], lf[TY], rt[TY], x[TY * TX];
#pragma unroll
    for (int ix = 0; ix < TX; ++ix) {
        up[ix] = *reinterpret_cast<const double *>(smem + nb_up + 8 * ((TY - 1) * TX + ix) + OFF_R);
        dn[ix] = *reinterpret_cast<const double *>(smem + nb_dn + 8 * ix + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy) {
        lf[iy] = *reinterpret_cast<const double *>(smem + nb_lf + 8 * (iy * TX + TX - 1) + OFF_R);
        rt[iy] = *reinterpret_cast<const double *>(smem + nb_rt + 8 * (iy * TX) + OFF_R);
    }
    double m = 0.0;
#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const double v_up = iy > 0 ? cur[c - TX] : up[ix];
            const double v_lf = ix > 0 ? cur[c - 1] : lf[iy];
            const double v_rt = ix < TX - 1 ? cur[c + 1] : rt[iy];
            const double v_dn = iy < TY - 1 ? cur[c + TX] : dn[ix];
            double acc = fma(w[c][0], v_up, 0.0);
            acc = fma(w[c][1], v_lf, acc);
            acc = fma(w[c][2], cur[c], acc);
            acc = fma(w[c][3], v_rt, acc);
            acc = fma(w[c][4], v_dn, acc);
            x[c] = acc;
        }
#pragma unroll
    for (int c = 0; c < TY * TX; ++c) {
        *reinterpret_cast<double *>(smem + own + 8 * c + OFF_W) = x[c];
        cur[c] = x[c];
        m = fmax(m, x[c]);
    }
    return m;
}

template <int TY, int TX, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) backward_grid5_kernel(const SuccBatch bt, const int n) {
    using Cfg = Grid5Cfg<TY, TX, MAXT>;
    constexpr int C = Cfg::C, K = 5, A = 4, STRIDE = Cfg::STRIDE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *scratch = reinterpret_cast<double *>(smem_raw + 2 * STRIDE);      // 32 doubles
    double *lin = reinterpret_cast<double *>(smem_raw + 2 * STRIDE + 256);    // S doubles: zs by state index

    SuccArgs a = bt.a;
    offset_succ(a, bt, blockIdx.x);
    const int S = a.S, tid = threadIdx.x;
    const int ntx = n / TX, nty = n / TY;
    const bool live = tid < ntx * nty;
    const int tx = live ? tid % ntx : 0, ty = live ? tid / ntx : 0;

    double w[C][5], cur[C];
    const uint32_t slot = 8u * Cfg::PITCH;
    const uint32_t own = slot * tid;
    const uint32_t nb_up = (live && ty > 0) ? own - slot * ntx : own;
    const uint32_t nb_dn = (live && ty < nty - 1) ? own + slot * ntx : own;
    const uint32_t nb_lf = (live && tx > 0) ? own - slot : own;
    const uint32_t nb_rt = (live && tx < ntx - 1) ? own + slot : own;

    double max_abs_r = 0.0;
#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const int s = (ty * TY + iy) * n + tx * TX + ix;
#pragma unroll
            for (int k = 0; k < 5; ++k) w[c][k] = 0.0;
            double z0 = 0.0;
            if (live) {
                const double r = a.reward[s];
                max_abs_r = fmax(max_abs_r, fabs(r));
                const double er = exp(r);                                   // np.exp(reward)   :142
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int succ = a.idx[(size_t)j * S + s];
                    double q = 0.0;
#pragma unroll
                    for (int aa = 0; aa < A; ++aa) q += __ldg(a.p + ((size_t)aa * K + j) * S + s);
                    q *= er;
                    const int off = succ - s;
                    w[c][0] += (off == -n) ? q : 0.0;
                    w[c][1] += (off == -1) ? q : 0.0;
                    w[c][2] += (off == 0) ? q : 0.0;
                    w[c][3] += (off == 1) ? q : 0.0;
                    w[c][4] += (off == n) ? q : 0.0;
                }
                z0 = a.term[s] ? 1.0 : 0.0;                                 // zs[terminal] = 1.0  :146-147
            }
            cur[c] = z0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c) = z0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c + STRIDE) = 0.0;
        }
    const int R = backward_rescale_period(block_max(max_abs_r, scratch), A);
    __syncthreads();

    // ---- n_sweeps - 1 merged-weight sweeps -------------------------------------------------
    const int n_lin = a.n_sweeps - 1;
    for (int t = 0; t < n_lin; ++t) {
        const double m = (t & 1)
            ? lin_grid5_sweep<TY, TX, MAXT, STRIDE, 0>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, cur)
            : lin_grid5_sweep<TY, TX, MAXT, 0, STRIDE>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, cur);
        if ((t + 1) % R == 0 && t + 1 < n_lin) {
            const double gm = block_max(m, scratch);           // two barriers inside
            if (gm > 0.0 && gm < INFINITY) {
                const int e = frexp_exponent(gm);
                const uint32_t offw = (t & 1) ? 0u : (uint32_t)STRIDE;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    cur[c] = ldexp(cur[c], -e);
                    *reinterpret_cast<double *>(smem_raw + own + 8 * c + offw) = cur[c];
                }
            }
        }
        __syncthreads();
    }

    // ---- last sweep, exactly as the reference evaluates it -----------------------------------
    if (live) {
#pragma unroll
        for (int iy = 0; iy < TY; ++iy)
#pragma unroll
            for (int ix = 0; ix < TX; ++ix) lin[(ty * TY + iy) * n + tx * TX + ix] = cur[iy * TX + ix];
    }
    __syncthreads();
    if (live && a.n_sweeps > 0) {
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            const int s = (ty * TY + c / TX) * n + tx * TX + c % TX;
            const double er = exp(a.reward[s]);
            double za[A];
            const double zs = succ_update<kOpBackward, 4>(
                A, K, [&](int aa, int j) { return __ldg(a.p + ((size_t)aa * K + j) * S + s); },
                [&](int j) { return lin[__ldg(a.idx + (size_t)j * S + s)]; }, er, 0.0, 0.0, 0, za);
#pragma unroll
            for (int aa = 0; aa < A; ++aa) a.policy[(size_t)s * A + aa] = za[aa] / zs;      // :159
        }
    }
}

// ---------------------------------------------------------------------------
// Cluster (push) variant of the stencil-tiled backward pass: ONE world spread over a thread-block
// cluster, for worlds too large for one CTA (BASELINE configs[2], 128 x 128: 2S = 32 768 partition
// sweeps, 1.7 us each behind a grid barrier).  Same merged-weight sweeps, same exact last sweep and
// same power-of-two rescale schedule as backward_grid5_kernel -- bitwise identical policies --
// with the exchange of svf_grid5_push_kernel: boundary rows travel by st.async into the
// neighbouring CTA's halo buffer, one mbarrier wait per sweep, no barrier.cluster in the loop.
// The sweep count is fixed (maxent.py:154), so there are no votes; on the rescale sweeps (every R)
// each warp sends its maximum to every CTA the same way, before the rows of that sweep are pushed.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void st_async_f64(uint32_t raddr, double v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];"
                 ::"r"(raddr), "d"(v), "r"(rbar) : "memory");
}

template <int TY, int TX, int MAXT>
struct BwdPushCfg {
    using G = Grid5Cfg<TY, TX, MAXT>;
    static constexpr int kMaxN = 128, kMaxCta = 16, NWARP = MAXT / 32;
    static constexpr int SLOTS = MAXT * G::PITCH * 8;
    static constexpr int HALO = 2 * kMaxN * 8;                            // [up row | down row]
    static constexpr int STRIDE = SLOTS + HALO;                           // bytes between the parities
    static constexpr int MAXTAB = 2 * STRIDE;                             // double[2][kMaxCta * NWARP]: warp maxima, by rescale parity
    static constexpr int MAXTAB_BYTES = kMaxCta * NWARP * 8;
    static constexpr int MBAR = MAXTAB + 2 * MAXTAB_BYTES;                // uint64[2] rows by parity, uint64[2] maxima by rescale parity
    static constexpr int SCRATCH = MBAR + 32;                             // 32 doubles + 16 doubles (cluster max of |r|)
    static constexpr int LIN = SCRATCH + 48 * 8;                          // (R TY + 2) n doubles: zs window of the last sweep
};

template <int TY, int TX, int MAXT, int P>
__device__ __forceinline__ bool bwd_push_iter(unsigned char *smem, const uint32_t sbase, const uint32_t own,
                                              const uint32_t nb_up, const uint32_t nb_dn, const uint32_t nb_lf,
                                              const uint32_t nb_rt, const uint32_t push_up, const uint32_t push_dn,
                                              const uint32_t bar_up, const uint32_t bar_dn, const uint32_t max_dst,
                                              const uint32_t max_bar, const uint32_t expect, const int ncta, const int nwarp,
                                              const double (&w)[TY * TX][5], double (&cur)[TY * TX], const int t,
                                              const bool rescale, int &n_rescale) {
    using Cfg = BwdPushCfg<TY, TX, MAXT>;
    constexpr int C = TY * TX, OFF_R = P * Cfg::STRIDE, OFF_W = (P ^ 1) * Cfg::STRIDE;
    if (threadIdx.x == 0) mbar_arrive_expect_tx(sbase + Cfg::MBAR + 8 * (P ^ 1), expect);
    if (t > 0 && !mbar_wait(sbase + Cfg::MBAR + 8 * P, ((t - 1) >> 1) & 1)) return false;
    double up[TX], dn[TX], lf[TY], rt[TY], x[C];
#pragma unroll
    for (int ix = 0; ix < TX; ++ix) {
        up[ix] = *reinterpret_cast<const double *>(smem + nb_up + 8 * ((TY - 1) * TX + ix) + OFF_R);
        dn[ix] = *reinterpret_cast<const double *>(smem + nb_dn + 8 * ix + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy) {
        lf[iy] = *reinterpret_cast<const double *>(smem + nb_lf + 8 * (iy * TX + TX - 1) + OFF_R);
        rt[iy] = *reinterpret_cast<const double *>(smem + nb_rt + 8 * (iy * TX) + OFF_R);
    }
#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const double v_up = iy > 0 ? cur[c - TX] : up[ix];
            const double v_lf = ix > 0 ? cur[c - 1] : lf[iy];
            const double v_rt = ix < TX - 1 ? cur[c + 1] : rt[iy];
            const double v_dn = iy < TY - 1 ? cur[c + TX] : dn[ix];
            double acc = fma(w[c][0], v_up, 0.0);
            acc = fma(w[c][1], v_lf, acc);
            acc = fma(w[c][2], cur[c], acc);
            acc = fma(w[c][3], v_rt, acc);
            acc = fma(w[c][4], v_dn, acc);
            x[c] = acc;
        }
    if (rescale) {
        // exact power-of-two rescale by the exponent of the cluster-wide maximum (range extension)
        const int rp = n_rescale & 1;
        const uint32_t mbar = sbase + Cfg::MBAR + 16 + 8 * rp;
        if (threadIdx.x == 0) mbar_arrive_expect_tx(mbar, 8u * ncta * nwarp);
        double m = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) m = fmax(m, x[c]);
        m = warp_max(m);
        if ((threadIdx.x & 31) < ncta) st_async_f64(max_dst + Cfg::MAXTAB_BYTES * rp, m, max_bar + 8 * rp);
        if (!mbar_wait(mbar, (n_rescale >> 1) & 1)) return false;
        const double *tab = reinterpret_cast<const double *>(smem + Cfg::MAXTAB + Cfg::MAXTAB_BYTES * rp);
        double gm = 0.0;
        for (int i = threadIdx.x & 31; i < ncta * nwarp; i += 32) {
            const double u = tab[i];
            gm = (u > gm || u != u) ? u : gm;
        }
        gm = warp_max(gm);
        if (gm > 0.0 && gm < INFINITY) {
            const int e = frexp_exponent(gm);
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = ldexp(x[c], -e);
        }
        ++n_rescale;
    }
    if (push_up) {
#pragma unroll
        for (int ix = 0; ix < TX; ix += 2) st_async_2f64(push_up + 8 * ix + OFF_W, x[ix], x[ix + 1], bar_up + 8 * (P ^ 1));
    }
    if (push_dn) {
#pragma unroll
        for (int ix = 0; ix < TX; ix += 2)
            st_async_2f64(push_dn + 8 * ix + OFF_W, x[(TY - 1) * TX + ix], x[(TY - 1) * TX + ix + 1], bar_dn + 8 * (P ^ 1));
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        *reinterpret_cast<double *>(smem + own + 8 * c + OFF_W) = x[c];
        cur[c] = x[c];
    }
    __syncthreads();
    return true;
}

template <int TY, int TX, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) backward_grid5_push_kernel(const SuccBatch bt, const int n, const int R) {
    using Cfg = BwdPushCfg<TY, TX, MAXT>;
    constexpr int C = TY * TX, K = 5, A = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cgx::cluster_group cl = cgx::this_cluster();
    const int ncta = (int)cl.num_blocks(), crank = (int)cl.block_rank();
    const uint32_t sbase = smem_u32(smem_raw);
    double *scratch = reinterpret_cast<double *>(smem_raw + Cfg::SCRATCH);
    double *rmax_tab = scratch + 32;                                      // [ncta]: every CTA's max |r|
    double *lin = reinterpret_cast<double *>(smem_raw + Cfg::LIN);

    SuccArgs a = bt.a;
    offset_succ(a, bt, blockIdx.x / ncta);
    const int S = a.S, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int ntx = n / TX;
    const bool live = tid < ntx * R;
    const int tx = live ? tid % ntx : 0, lty = live ? tid / ntx : 0;
    const int gty = crank * R + lty, nty = n / TY;

    double w[C][5], cur[C];
    const uint32_t slot = 8u * Cfg::G::PITCH;
    const uint32_t own = slot * tid;
    const uint32_t nb_lf = (live && tx > 0) ? own - slot : own;
    const uint32_t nb_rt = (live && tx < ntx - 1) ? own + slot : own;
    uint32_t nb_up = own, nb_dn = own, push_up = 0, push_dn = 0, bar_up = 0, bar_dn = 0;
    if (live && lty > 0) nb_up = own - slot * ntx;
    else if (live && gty > 0) {
        nb_up = Cfg::SLOTS + 8u * (tx * TX) - 8u * ((TY - 1) * TX);
        push_up = mapa_u32(sbase + Cfg::SLOTS + 8u * (Cfg::kMaxN + tx * TX), crank - 1);
        bar_up = mapa_u32(sbase + Cfg::MBAR, crank - 1);
    }
    if (live && lty < R - 1) nb_dn = own + slot * ntx;
    else if (live && gty < nty - 1) {
        nb_dn = Cfg::SLOTS + 8u * (Cfg::kMaxN + tx * TX);
        push_dn = mapa_u32(sbase + Cfg::SLOTS + 8u * (tx * TX), crank + 1);
        bar_dn = mapa_u32(sbase + Cfg::MBAR, crank + 1);
    }
    const uint32_t max_dst = lane < ncta ? mapa_u32(sbase + Cfg::MAXTAB + 8u * (crank * nwarp + warp), lane) : 0;
    const uint32_t max_bar = lane < ncta ? mapa_u32(sbase + Cfg::MBAR + 16, lane) : 0;
    const uint32_t expect = (crank > 0 ? 8u * n : 0u) + (crank < ncta - 1 ? 8u * n : 0u);

    // zero the halo rows and the maxima tables before anybody writes into them
    for (int i = tid; i < Cfg::HALO / 8; i += blockDim.x) {
        *reinterpret_cast<double *>(smem_raw + Cfg::SLOTS + 8 * i) = 0.0;
        *reinterpret_cast<double *>(smem_raw + Cfg::SLOTS + 8 * i + Cfg::STRIDE) = 0.0;
    }
    for (int i = tid; i < 2 * Cfg::MAXTAB_BYTES / 8; i += blockDim.x)
        *reinterpret_cast<double *>(smem_raw + Cfg::MAXTAB + 8 * i) = 0.0;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(sbase + Cfg::MBAR + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cl.sync();

    double max_abs_r = 0.0;
#pragma unroll
    for (int iy = 0; iy < TY; ++iy)
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) {
            const int c = iy * TX + ix;
            const int s = (gty * TY + iy) * n + tx * TX + ix;
#pragma unroll
            for (int k = 0; k < 5; ++k) w[c][k] = 0.0;
            double z0 = 0.0;
            if (live) {
                const double r = a.reward[s];
                max_abs_r = fmax(max_abs_r, fabs(r));
                const double er = exp(r);                                   // np.exp(reward)   :142
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int succ = a.idx[(size_t)j * S + s];
                    double q = 0.0;
#pragma unroll
                    for (int aa = 0; aa < A; ++aa) q += __ldg(a.p + ((size_t)aa * K + j) * S + s);
                    q *= er;
                    const int off = succ - s;
                    w[c][0] += (off == -n) ? q : 0.0;
                    w[c][1] += (off == -1) ? q : 0.0;
                    w[c][2] += (off == 0) ? q : 0.0;
                    w[c][3] += (off == 1) ? q : 0.0;
                    w[c][4] += (off == n) ? q : 0.0;
                }
                z0 = a.term[s] ? 1.0 : 0.0;                                 // zs[terminal] = 1.0  :146-147
            }
            cur[c] = z0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c) = z0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c + Cfg::STRIDE) = 0.0;
        }
    // the start vector's boundary rows and every CTA's max |r|: plain DSMEM stores, fenced by barrier.cluster
    if (push_up) {
        double *dst = reinterpret_cast<double *>(cl.map_shared_rank(smem_raw, crank - 1) + Cfg::SLOTS + 8 * (Cfg::kMaxN + tx * TX));
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) dst[ix] = cur[ix];
    }
    if (push_dn) {
        double *dst = reinterpret_cast<double *>(cl.map_shared_rank(smem_raw, crank + 1) + Cfg::SLOTS + 8 * (tx * TX));
#pragma unroll
        for (int ix = 0; ix < TX; ++ix) dst[ix] = cur[(TY - 1) * TX + ix];
    }
    const double cta_max_r = block_max(max_abs_r, scratch);
    if (tid < ncta) *cl.map_shared_rank(rmax_tab + crank, tid) = cta_max_r;
    cl.sync();
    double gmax_r = 0.0;
    for (int i = 0; i < ncta; ++i) {
        const double u = rmax_tab[i];
        gmax_r = (u > gmax_r || u != u) ? u : gmax_r;
    }
    const int RP = backward_rescale_period(gmax_r, A);

    // ---- n_sweeps - 1 merged-weight sweeps -------------------------------------------------
    const int n_lin = a.n_sweeps - 1;
    int n_rescale = 0;
    bool ok = true;
    for (int t = 0; t < n_lin && ok; t += 2) {
        ok = bwd_push_iter<TY, TX, MAXT, 0>(smem_raw, sbase, own, nb_up, nb_dn, nb_lf, nb_rt, push_up, push_dn, bar_up, bar_dn,
                                            max_dst, max_bar, expect, ncta, nwarp, w, cur, t,
                                            (t + 1) % RP == 0 && t + 1 < n_lin, n_rescale);
        if (!ok || t + 1 >= n_lin) break;
        ok = bwd_push_iter<TY, TX, MAXT, 1>(smem_raw, sbase, own, nb_up, nb_dn, nb_lf, nb_rt, push_up, push_dn, bar_up, bar_dn,
                                            max_dst, max_bar, expect, ncta, nwarp, w, cur, t + 1,
                                            (t + 2) % RP == 0 && t + 2 < n_lin, n_rescale);
    }
    if (!ok) {
        // a peer CTA never arrived (4 s timeout): make the failure loud instead of returning a half-swept policy
        if (live)
            for (int c = 0; c < C; ++c) {
                const int s = (gty * TY + c / TX) * n + tx * TX + c % TX;
                for (int aa = 0; aa < A; ++aa) a.policy[(size_t)s * A + aa] = __longlong_as_double(0x7ff8000000000000ll);
            }
    } else {
        // ---- last sweep, exactly as the reference evaluates it ---------------------------------
        const int pf = n_lin > 0 ? (n_lin & 1) : 0;                       // parity holding the final rows
        if (n_lin > 0) ok = mbar_wait(sbase + Cfg::MBAR + 8 * pf, ((n_lin - 1) >> 1) & 1);
        const int row0 = crank * R * TY;                                  // first grid row of this CTA; window starts one above
        if (live) {
#pragma unroll
            for (int iy = 0; iy < TY; ++iy)
#pragma unroll
                for (int ix = 0; ix < TX; ++ix) lin[(lty * TY + iy + 1) * n + tx * TX + ix] = cur[iy * TX + ix];
        }
        for (int i = tid; i < n; i += blockDim.x) {
            lin[i] = *reinterpret_cast<const double *>(smem_raw + Cfg::SLOTS + 8 * i + pf * Cfg::STRIDE);
            lin[(R * TY + 1) * n + i] = *reinterpret_cast<const double *>(smem_raw + Cfg::SLOTS + 8 * (Cfg::kMaxN + i) + pf * Cfg::STRIDE);
        }
        __syncthreads();
        if (live && a.n_sweeps > 0) {
            const int base = (row0 - 1) * n;
#pragma unroll 1
            for (int c = 0; c < C; ++c) {
                const int s = (gty * TY + c / TX) * n + tx * TX + c % TX;
                const double er = exp(a.reward[s]);
                double za[A];
                const double zs = succ_update<kOpBackward, 4>(
                    A, K, [&](int aa, int j) { return __ldg(a.p + ((size_t)aa * K + j) * S + s); },
                    [&](int j) { return lin[__ldg(a.idx + (size_t)j * S + s) - base]; }, er, 0.0, 0.0, 0, za);
#pragma unroll
                for (int aa = 0; aa < A; ++aa) a.policy[(size_t)s * A + aa] = za[aa] / zs;      // :159
            }
        }
    }
    cl.sync();
}

template <int TY, int TX, int MAXT, int MINB>
static int launch_backward_grid5(const SuccBatch &bt, int B, int n, cudaStream_t st);

template <class Kern> static int prep_smem(Kern k, size_t bytes);
static inline int round_up32(int x);
template <int TY, int TX, int MAXT, int MINB>
static int launch_svf_grid5(const SvfBatch &bt, int B, int n, cudaStream_t st) {
    using Cfg = Grid5Cfg<TY, TX, MAXT>;
    auto k = svf_grid5_kernel<TY, TX, MAXT, MINB>;
    const size_t sm = 2 * (size_t)Cfg::STRIDE + 16;
    if (int rc = prep_smem(k, sm)) return rc;
    const int threads = round_up32((n / TX) * (n / TY));
    k<<<B, threads, sm, st>>>(bt, n);
    return IRLB200_OK;
}

// ---------------------------------------------------------------------------
// grid (cooperative) kernels -- one problem
// ---------------------------------------------------------------------------
struct GridWork {
    double *buf0, *buf1;
    GridSyncState *gs;
};

__device__ __forceinline__ void carve_grid(GridTopo &tp, const GridWork &w) {
    __shared__ double s_scratch[32];
    __shared__ unsigned long long s_word;
    __shared__ int s_flag;
    tp.buf0 = w.buf0;
    tp.buf1 = w.buf1;
    tp.gs = w.gs;
    tp.seq = 0;
    tp.scratch = s_scratch;
    tp.s_word = &s_word;
    tp.flag = &s_flag;
}

template <int OP, int A_T, int K_T, int SPT_T, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
    succ_grid_kernel(const SuccArgs a, const GridWork w, int32_t *n_iter, int32_t *status) {
    GridTopo tp;
    carve_grid(tp, w);
    succ_phase<GridTopo, OP, A_T, K_T, SPT_T>(tp, a, n_iter, status);
}

template <int A_T, int K_T, int SPT_T, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
    svf_grid_kernel(const SvfArgs a, const GridWork w, int32_t *n_iter, int32_t *status) {
    GridTopo tp;
    carve_grid(tp, w);
    svf_phase<GridTopo, A_T, K_T, SPT_T>(tp, a, n_iter, status);
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
template <class Kern>
static int prep_smem(Kern k, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(smem)");
    }
    return IRLB200_OK;
}

static inline int round_up32(int x) { return (x + 31) & ~31; }

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

#define LAUNCH_CHECK(what)                                     \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return fail_cuda(e__, what);   \
    } while (0)

// fast path = the table shape of every 4-action grid world
static inline bool is_fast_shape(int A, int K) { return A == 4 && K == 5; }

template <int TY, int TX, int MAXT, int MINB>
static int launch_backward_grid5(const SuccBatch &bt, int B, int n, cudaStream_t st) {
    using Cfg = Grid5Cfg<TY, TX, MAXT>;
    auto k = backward_grid5_kernel<TY, TX, MAXT, MINB>;
    const size_t sm = 2 * (size_t)Cfg::STRIDE + 256 + sizeof(double) * (size_t)n * n;
    if (int rc = prep_smem(k, sm)) return rc;
    k<<<B, round_up32((n / TX) * (n / TY)), sm, st>>>(bt, n);
    return IRLB200_OK;
}

// cluster launch of the tiled forward pass: B worlds, `ncta` CTAs each
static int launch_svf_grid5_cluster(const SvfBatch &bt, int B, int n, int ncta, cudaStream_t st) {
    using Cfg = Grid5Cfg<2, 4, 256>;
    auto k = svf_grid5_cluster_kernel<2, 4, 256>;
    const int ntx = n / 4, nty = n / 2, R = nty / ncta;
    const size_t sm = 2 * (size_t)Cfg::STRIDE + 32;
    if (int rc = prep_smem(k, sm)) return rc;
    if (ncta > 8) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(non-portable cluster)");
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B * ncta));
    cfg.blockDim = dim3((unsigned)round_up32(ntx * R));
    cfg.dynamicSmemBytes = sm;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ncta;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, bt, n, R);
    if (e != cudaSuccess) return fail_cuda(e, "cudaLaunchKernelEx(svf cluster)");
    return IRLB200_OK;
}

// push variant (st.async + mbarrier, no barrier.cluster per sweep): the default cluster kernel
static int launch_svf_grid5_push(const SvfBatch &bt, int B, int n, int ncta, cudaStream_t st) {
    using Cfg = PushCfg<2, 4, 256>;
    auto k = svf_grid5_push_kernel<2, 4, 256>;
    const int ntx = n / 4, nty = n / 2, R = nty / ncta;
    if (n > Cfg::kMaxN) return fail(IRLB200_ELIMIT, "cluster mode: grid row wider than 128 cells");
    if (int rc = prep_smem(k, (size_t)Cfg::BYTES)) return rc;
    if (ncta > 8) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(non-portable cluster)");
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B * ncta));
    cfg.blockDim = dim3((unsigned)round_up32(ntx * R));
    cfg.dynamicSmemBytes = (size_t)Cfg::BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ncta;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, bt, n, R);
    if (e != cudaSuccess) return fail_cuda(e, "cudaLaunchKernelEx(svf cluster push)");
    return IRLB200_OK;
}

// cluster launch of the tiled backward pass (push exchange)
static int launch_backward_grid5_push(const SuccBatch &bt, int B, int n, int ncta, cudaStream_t st) {
    using Cfg = BwdPushCfg<2, 4, 256>;
    auto k = backward_grid5_push_kernel<2, 4, 256>;
    const int ntx = n / 4, nty = n / 2, R = nty / ncta;
    if (n > Cfg::kMaxN) return fail(IRLB200_ELIMIT, "cluster mode: grid row wider than 128 cells");
    const size_t sm = (size_t)Cfg::LIN + sizeof(double) * (size_t)(R * 2 + 2) * n;
    if (int rc = prep_smem(k, sm)) return rc;
    if (ncta > 8) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(non-portable cluster)");
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B * ncta));
    cfg.blockDim = dim3((unsigned)round_up32(ntx * R));
    cfg.dynamicSmemBytes = sm;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ncta;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, bt, n, R);
    if (e != cudaSuccess) return fail_cuda(e, "cudaLaunchKernelEx(backward cluster push)");
    return IRLB200_OK;
}

static int launch_svf_cluster(const SvfBatch &bt, int B, int n, int ncta, cudaStream_t st) {
    return env_int("IRLB200_CLUSTER_PUSH", 1) ? launch_svf_grid5_push(bt, B, n, ncta, st)
                                              : launch_svf_grid5_cluster(bt, B, n, ncta, st);
}

// pick a cluster size for an n x n world: tile rows must split evenly, <= 256 threads per CTA.
// Push kernel (default): the flight time does not grow with the cluster, so the LARGEST portable
// cluster that still leaves a warp of tiles per CTA wins (measured, us per sweep: 64 x 64 -- 2 CTAs
// 0.45, 4: 0.37, 8: 0.36, 16: 0.40; 128 x 128 -- 8: 0.53, 16: 0.60); barrier.cluster kernel
// (IRLB200_CLUSTER_PUSH=0): the barrier grows with the cluster, the SMALLEST admissible size wins
// (128 x 128 -- 8: 0.95, 16: 1.20).  IRLB200_CLUSTER_SIZE forces a size; `max_c` caps it.
static int cluster_size_for(int n, int max_c = 16, bool force_push = false) {
    if (n % 4 || n % 2) return 0;
    const int ntx = n / 4, nty = n / 2;
    const int pref = env_int("IRLB200_CLUSTER_SIZE", 0);
    const bool push = force_push || env_int("IRLB200_CLUSTER_PUSH", 1) != 0;
    int smallest = 0, best = 0;
    for (int c : {2, 4, 8, 16}) {
        if (c > max_c || nty % c) continue;
        const int threads = ntx * (nty / c);
        if (threads > 256) continue;
        if (pref > 0) {
            if (c == pref) return c;
            continue;
        }
        if (!smallest) smallest = c;
        if (c <= 8 && threads >= 32) best = c;
    }
    return (push && best) ? best : smallest;
}

// ---- CTA: successor phases --------------------------------------------------
template <int OP>
static int launch_succ_cta(const SuccBatch &bt, int B, cudaStream_t st) {
    const int S = bt.a.S, A = bt.a.A, K = bt.a.K;
    const size_t smem = cta_smem_bytes(S, A, false);
    const bool fast = is_fast_shape(A, K);
    const int force_stream = env_int("IRLB200_FORCE_STREAMED", 0);
    const int gn = bt.a.grid_n;
    if (OP == kOpBackward && fast && !force_stream && gn > 0 && gn * gn == S && gn % 4 == 0 &&
        env_int("IRLB200_BWD_TILE", 1) && bt.a.n_sweeps > 0) {
        int rc = -100;
        if ((gn / 4) * (gn / 2) <= 128) rc = launch_backward_grid5<2, 4, 128, 3>(bt, B, gn, st);
        else if ((gn / 4) * (gn / 2) <= 512) rc = launch_backward_grid5<2, 4, 512, 1>(bt, B, gn, st);
        if (rc != -100) {
            if (rc) return rc;
            LAUNCH_CHECK("backward_grid5_kernel");
            return IRLB200_OK;
        }
    }
    // soft-VI keeps 4 exp + 1 log of temporaries live: two states per thread spill at 128 registers, and
    // the streamed kernel (rows re-read from L1 each sweep, 64 registers) measured 1.6x faster on
    // 1 024-state worlds, so only the other operators take the two-states-per-thread variant
    const bool reg2 = fast && S > 512 && S <= 1024 && !force_stream &&
                      (OP != kOpSoftVI || env_int("IRLB200_SOFTVI_REG2", 0));
    if (fast && S <= 512 && !force_stream) {
        auto k = succ_cta_kernel<OP, 4, 5, 1, 512, 1>;
        if (int rc = prep_smem(k, smem)) return rc;
        k<<<B, round_up32(S), smem, st>>>(bt);
    } else if (reg2) {
        auto k = succ_cta_kernel<OP, 4, 5, 2, 512, 1>;
        if (int rc = prep_smem(k, smem)) return rc;
        k<<<B, round_up32((S + 1) / 2), smem, st>>>(bt);
    } else if (fast) {
        auto k = succ_cta_kernel<OP, 4, 5, 0, 1024, 1>;
        if (int rc = prep_smem(k, smem)) return rc;
        const int cap = env_int("IRLB200_STREAM_THREADS", 1024);
        k<<<B, S < cap ? round_up32(S) : cap, smem, st>>>(bt);
    } else {
        if (A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
        auto k = succ_cta_kernel<OP, 0, 0, 0, 1024, 1>;
        if (int rc = prep_smem(k, smem)) return rc;
        k<<<B, S < 1024 ? round_up32(S) : 1024, smem, st>>>(bt);
    }
    LAUNCH_CHECK("succ_cta_kernel");
    return IRLB200_OK;
}

// ---- CTA: forward phase -----------------------------------------------------
static int launch_svf_cta(SvfBatch bt, int B, cudaStream_t st) {
    const int S = bt.a.S, A = bt.a.A, K = bt.a.K;
    const size_t smem = cta_smem_bytes(S, A, false);
    const bool fast = is_fast_shape(A, K);
    const int force_stream = env_int("IRLB200_FORCE_STREAMED", 0);
    const int spt_pref = env_int("IRLB200_SVF_SPT", 0);
    const int gn = bt.a.grid_n;
    const int tile = env_int("IRLB200_SVF_TILE", 24);       // TY*10 + TX; 0 disables the tiled kernels
    if (fast && !force_stream && gn > 0 && tile > 0 && gn * gn == S) {
        bt.a.w_scratch = nullptr;
        int rc = -100;
        if (tile == 24 && gn % 4 == 0 && gn % 2 == 0 && (gn / 4) * (gn / 2) <= 128) rc = launch_svf_grid5<2, 4, 128, 3>(bt, B, gn, st);
        else if (tile == 124 && gn % 4 == 0 && (gn / 4) * (gn / 2) <= 128) rc = launch_svf_grid5<2, 4, 128, 4>(bt, B, gn, st);
        else if (tile == 142 && gn % 4 == 0 && (gn / 2) * (gn / 4) <= 128) rc = launch_svf_grid5<4, 2, 128, 4>(bt, B, gn, st);
        else if (tile == 24 && gn % 4 == 0 && (gn / 4) * (gn / 2) <= 512) rc = launch_svf_grid5<2, 4, 512, 1>(bt, B, gn, st);
        else if (tile == 44 && gn % 4 == 0 && (gn / 4) * (gn / 4) <= 64) rc = launch_svf_grid5<4, 4, 64, 4>(bt, B, gn, st);
        else if (tile == 44 && gn % 4 == 0 && (gn / 4) * (gn / 4) <= 256) rc = launch_svf_grid5<4, 4, 256, 1>(bt, B, gn, st);
        else if (tile == 22 && gn % 2 == 0 && (gn / 2) * (gn / 2) <= 256) rc = launch_svf_grid5<2, 2, 256, 2>(bt, B, gn, st);
        else if (tile == 42 && gn % 4 == 0 && (gn / 2) * (gn / 4) <= 128) rc = launch_svf_grid5<4, 2, 128, 3>(bt, B, gn, st);
        if (rc != -100) {
            if (rc) return rc;
            LAUNCH_CHECK("svf_grid5_kernel");
            return IRLB200_OK;
        }
    }
    if (fast && !force_stream && S <= 4096) {
        // hand-tuned kernel; (states per thread, buffer stride) picked by size.  The block is
        // ceil(S / SPT) threads rounded to a warp; padded states carry zero weights.
        int spt = spt_pref ? spt_pref : (S <= 128 ? 1 : (S <= 2048 ? 2 : 4));
        bt.a.w_scratch = nullptr;
#define SVF_FAST(SPT_, STRIDE_, MAXT_, MINB_)                                               \
    do {                                                                                    \
        auto k = svf_cta_fast_kernel<SPT_, STRIDE_, MAXT_, MINB_>;                          \
        const size_t sm = 2 * (size_t)(STRIDE_) + 16;                                       \
        if (int rc = prep_smem(k, sm)) return rc;                                           \
        k<<<B, round_up32((S + (SPT_) - 1) / (SPT_)), sm, st>>>(bt);                        \
    } while (0)
        if (S <= 256) {
            if (spt == 1) SVF_FAST(1, 2048, 256, 1); else if (spt == 2) SVF_FAST(2, 2048, 128, 1); else SVF_FAST(4, 2048, 64, 1);
        } else if (S <= 1024) {
            if (spt == 1) SVF_FAST(1, 8192, 1024, 1); else if (spt == 2) SVF_FAST(2, 8192, 512, 2); else SVF_FAST(4, 8192, 256, 2);
        } else {
            if (spt <= 2 && S <= 2048) SVF_FAST(2, 32768, 1024, 1); else SVF_FAST(4, 32768, 1024, 1);
        }
#undef SVF_FAST
    } else {
        if (A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
        double *ws = nullptr;
        if (int rc = workspace(0, (size_t)B * S * K * sizeof(double), (void **)&ws)) return rc;
        bt.a.w_scratch = ws;
        if (fast) {
            auto k = svf_cta_kernel<4, 5, 0, 1024, 1>;
            if (int rc = prep_smem(k, smem)) return rc;
            k<<<B, S < 1024 ? round_up32(S) : 1024, smem, st>>>(bt);
        } else {
            auto k = svf_cta_kernel<0, 0, 0, 1024, 1>;
            if (int rc = prep_smem(k, smem)) return rc;
            k<<<B, S < 1024 ? round_up32(S) : 1024, smem, st>>>(bt);
        }
    }
    LAUNCH_CHECK("svf_cta_kernel");
    return IRLB200_OK;
}

// ---- CTA: fused step ----------------------------------------------------------
template <bool CAUSAL>
static int launch_step_cta(StepBatch bt, int B, cudaStream_t st) {
    const int S = bt.s.S, A = bt.s.A;
    const size_t smem = cta_smem_bytes(S, A, true);
    const bool fast = is_fast_shape(A, bt.s.K) && is_fast_shape(A, bt.f.K);
    const int force_stream = env_int("IRLB200_FORCE_STREAMED", 0);
    if (fast && S <= 32 && !force_stream && env_int("IRLB200_WARP_STEP", 1)) {
        // tiny worlds: one warp per world, four worlds per CTA, no barriers
        bt.f.w_scratch = nullptr;
        step_warp_kernel<CAUSAL><<<(B + 3) / 4, 128, 0, st>>>(bt, B);
        LAUNCH_CHECK("step_warp_kernel");
        return IRLB200_OK;
    }
    if (fast && S <= 512 && !force_stream) {
        bt.f.w_scratch = nullptr;
        auto k = step_cta_kernel<CAUSAL, 4, 5, 1, 512>;
        if (int rc = prep_smem(k, smem)) return rc;
        k<<<B, round_up32(S), smem, st>>>(bt);
    } else if (fast && S <= 1024 && !force_stream) {
        bt.f.w_scratch = nullptr;
        auto k = step_cta_kernel<CAUSAL, 4, 5, 2, 512>;
        if (int rc = prep_smem(k, smem)) return rc;
        k<<<B, round_up32((S + 1) / 2), smem, st>>>(bt);
    } else {
        if (A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
        if (bt.s.K != bt.f.K && fast) return fail(IRLB200_EINVAL, "internal: fused fast shape mismatch");
        double *ws = nullptr;
        if (int rc = workspace(0, (size_t)B * S * bt.f.K * sizeof(double), (void **)&ws)) return rc;
        bt.f.w_scratch = ws;
        if (fast) {
            auto k = step_cta_kernel<CAUSAL, 4, 5, 0, 1024>;
            if (int rc = prep_smem(k, smem)) return rc;
            k<<<B, S < 1024 ? round_up32(S) : 1024, smem, st>>>(bt);
        } else {
            // run-time shapes: the two phases may have different K; the phase
            // templates read K from their own argument block when K_T == 0
            auto k = step_cta_kernel<CAUSAL, 0, 0, 0, 1024>;
            if (int rc = prep_smem(k, smem)) return rc;
            k<<<B, S < 1024 ? round_up32(S) : 1024, smem, st>>>(bt);
        }
    }
    LAUNCH_CHECK("step_cta_kernel");
    return IRLB200_OK;
}

// ---- grid ---------------------------------------------------------------------
struct GridPlan {
    int blocks, threads;
};

template <class Kern>
static int plan_grid(Kern k, int S, int threads, int states_per_thread, GridPlan *out) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, 0);
    if (e != cudaSuccess) return fail_cuda(e, "occupancy");
    if (per_sm < 1) return fail(IRLB200_ELIMIT, "grid kernel does not fit on an SM");
    long long want = ((long long)S + (long long)threads * states_per_thread - 1) /
                     ((long long)threads * states_per_thread);
    long long cap = (long long)sms * per_sm;
    int max_blocks = env_int("IRLB200_GRID_BLOCKS", 0);
    if (max_blocks > 0 && max_blocks < cap) cap = max_blocks;
    out->blocks = (int)(want < cap ? want : cap);
    if (out->blocks < 1) out->blocks = 1;
    out->threads = threads;
    return IRLB200_OK;
}

static int grid_work(int S, GridWork *w, cudaStream_t st) {
    // workspace slot 1: [GridSyncState | buf0 | buf1]
    const size_t hdr = 256;
    unsigned char *base = nullptr;
    if (int rc = workspace(1, hdr + 2 * (size_t)S * sizeof(double), (void **)&base)) return rc;
    w->gs = reinterpret_cast<GridSyncState *>(base);
    w->buf0 = reinterpret_cast<double *>(base + hdr);
    w->buf1 = w->buf0 + S;
    cudaError_t e = cudaMemsetAsync(base, 0, hdr, st);
    if (e != cudaSuccess) return fail_cuda(e, "memset(grid sync)");
    return IRLB200_OK;
}

template <class Kern, class Args>
static int launch_coop(Kern k, const GridPlan &pl, const Args &a, const GridWork &w, int32_t *n_iter,
                       int32_t *status, cudaStream_t st) {
    void *params[] = {(void *)&a, (void *)&w, (void *)&n_iter, (void *)&status};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k, dim3(pl.blocks), dim3(pl.threads), params, 0, st);
    if (e != cudaSuccess) return fail_cuda(e, "cudaLaunchCooperativeKernel");
    return IRLB200_OK;
}

template <int OP>
static int launch_succ_grid(const SuccArgs &a, int32_t *n_iter, int32_t *status, cudaStream_t st) {
    GridWork w;
    if (int rc = grid_work(a.S, &w, st)) return rc;
    GridPlan pl;
    const bool fast = is_fast_shape(a.A, a.K);
    const int threads = env_int("IRLB200_GRID_THREADS", 256);
    if (fast) {
        // register-resident rows when one state per thread covers the problem
        auto kr = succ_grid_kernel<OP, 4, 5, 1, 256, 1>;
        GridPlan pr;
        if (int rc = plan_grid(kr, a.S, 256, 1, &pr)) return rc;
        if ((long long)pr.blocks * 256 >= a.S && !env_int("IRLB200_FORCE_STREAMED", 0))
            return launch_coop(kr, pr, a, w, n_iter, status, st);
        auto k = succ_grid_kernel<OP, 4, 5, 0, 512, 1>;
        if (int rc = plan_grid(k, a.S, threads, 1, &pl)) return rc;
        return launch_coop(k, pl, a, w, n_iter, status, st);
    }
    if (a.A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
    auto k = succ_grid_kernel<OP, 0, 0, 0, 512, 1>;
    if (int rc = plan_grid(k, a.S, threads, 1, &pl)) return rc;
    return launch_coop(k, pl, a, w, n_iter, status, st);
}

static int launch_svf_grid(SvfArgs a, int32_t *n_iter, int32_t *status, cudaStream_t st) {
    GridWork w;
    if (int rc = grid_work(a.S, &w, st)) return rc;
    GridPlan pl;
    const bool fast = is_fast_shape(a.A, a.K);
    const int threads = env_int("IRLB200_GRID_THREADS", 256);
    if (fast) {
        auto kr = svf_grid_kernel<4, 5, 1, 256, 1>;
        GridPlan pr;
        if (int rc = plan_grid(kr, a.S, 256, 1, &pr)) return rc;
        if ((long long)pr.blocks * 256 >= a.S && !env_int("IRLB200_FORCE_STREAMED", 0)) {
            a.w_scratch = nullptr;
            return launch_coop(kr, pr, a, w, n_iter, status, st);
        }
    }
    double *ws = nullptr;
    if (int rc = workspace(0, (size_t)a.S * a.K * sizeof(double), (void **)&ws)) return rc;
    a.w_scratch = ws;
    if (fast) {
        auto k = svf_grid_kernel<4, 5, 0, 512, 1>;
        if (int rc = plan_grid(k, a.S, threads, 1, &pl)) return rc;
        return launch_coop(k, pl, a, w, n_iter, status, st);
    }
    if (a.A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
    auto k = svf_grid_kernel<0, 0, 0, 512, 1>;
    if (int rc = plan_grid(k, a.S, threads, 1, &pl)) return rc;
    return launch_coop(k, pl, a, w, n_iter, status, st);
}

}  // namespace irlb200

// ===========================================================================
// C ABI
// ===========================================================================
using namespace irlb200;

static int max_states_cta_impl() {
    int dev = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // fused step needs (2 + A) * S + 34 doubles; quote the A = 4 figure
    long long s = ((long long)optin / 8 - 34) / 6;
    return (int)(s < 0 ? 0 : s);
}

extern "C" int irlb200_max_states_cta(void) { return max_states_cta_impl(); }
// cluster mode exists for the forward pass of grid-stencil tables: 16 CTAs x 256 tiles x 8 cells
extern "C" int irlb200_max_states_cluster(void) { return 16 * 256 * 8; }

static int check_tables(const irlb200_tables *t, bool need_succ, bool need_pred) {
    if (!t) return fail(IRLB200_EINVAL, "tables == NULL");
    if (t->S <= 0 || t->A <= 0) return fail(IRLB200_EINVAL, "S and A must be positive");
    if (need_succ && (!t->succ_idx || !t->succ_p || t->Ks <= 0)) return fail(IRLB200_EINVAL, "successor table missing");
    if (need_pred && (!t->pred_idx || !t->pred_p || t->Kp <= 0)) return fail(IRLB200_EINVAL, "predecessor table missing");
    return IRLB200_OK;
}

static int pick_mode(int mode, int B, int S, int A, bool fused, int *out) {
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const bool fits_cta = cta_smem_bytes(S, A, fused) <= (size_t)optin;
    if (mode == IRLB200_MODE_AUTO) mode = (fits_cta && (B > 1 || S <= 2048)) ? IRLB200_MODE_CTA : IRLB200_MODE_GRID;
    if (mode == IRLB200_MODE_CLUSTER) return fail(IRLB200_ELIMIT, "cluster mode is not available in this build");
    if (mode == IRLB200_MODE_CTA && !fits_cta) return fail(IRLB200_ELIMIT, "problem does not fit one CTA's shared memory");
    if (mode == IRLB200_MODE_GRID && B != 1) return fail(IRLB200_ELIMIT, "grid mode takes one problem per call");
    if (mode != IRLB200_MODE_CTA && mode != IRLB200_MODE_GRID) return fail(IRLB200_EINVAL, "unknown mode");
    *out = mode;
    return IRLB200_OK;
}

static void fill_succ(SuccArgs &a, const irlb200_tables *t) {
    a = SuccArgs{};
    a.S = t->S; a.A = t->A; a.K = t->Ks;
    a.idx = t->succ_idx; a.p = t->succ_p;
    a.grid_n = t->stencil_n;
}

static void succ_strides(SuccBatch &bt, const irlb200_tables *t) {
    bt.tab_idx_stride = t->shared ? 0 : (size_t)t->Ks * t->S;
    bt.tab_p_stride = t->shared ? 0 : (size_t)t->A * t->Ks * t->S;
}

extern "C" int irlb200_backward(const irlb200_tables *t, int B, const double *reward,
                                const uint8_t *terminal_mask, int mask_shared, int n_sweeps,
                                double *policy, int mode, void *stream) {
    if (int rc = check_tables(t, true, false)) return rc;
    if (B <= 0 || !reward || !terminal_mask || !policy || n_sweeps < 0) return fail(IRLB200_EINVAL, "backward: bad argument");
    // thread-block-cluster mode (push exchange): grid-stencil tables too large for the one-CTA tiled kernel
    const int cl_size = (t->stencil_n > 0 && t->stencil_n * t->stencil_n == t->S && t->A == 4 && t->Ks == 5 &&
                         t->stencil_n <= 128 && n_sweeps > 0)
                            ? cluster_size_for(t->stencil_n, 16, true) : 0;
    const bool want_cluster = mode == IRLB200_MODE_CLUSTER ||
                              (mode == IRLB200_MODE_AUTO && cl_size > 0 && t->stencil_n > 64 && B <= 64 &&
                               env_int("IRLB200_BWD_TILE", 1));
    if (mode == IRLB200_MODE_CLUSTER && cl_size == 0)
        return fail(IRLB200_ELIMIT, "cluster mode needs grid-stencil tables with n <= 128, n % 4 == 0");
    if (!want_cluster)
        if (int rc = pick_mode(mode, B, t->S, t->A, false, &mode)) return rc;
    SuccBatch bt{};
    fill_succ(bt.a, t);
    bt.a.reward = reward; bt.a.term = terminal_mask; bt.a.n_sweeps = n_sweeps; bt.a.policy = policy;
    succ_strides(bt, t);
    bt.term_stride = mask_shared ? 0 : (size_t)t->S;
    if (want_cluster) {
        if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
        int rc = launch_backward_grid5_push(bt, B, t->stencil_n, cl_size, (cudaStream_t)stream);
        if (rc != IRLB200_OK && cl_size > 8) {
            cudaGetLastError();
            const int c8 = cluster_size_for(t->stencil_n, 8, true);
            if (c8 > 0) rc = launch_backward_grid5_push(bt, B, t->stencil_n, c8, (cudaStream_t)stream);
        }
        if (rc == IRLB200_OK || mode == IRLB200_MODE_CLUSTER) return rc;
        cudaGetLastError();                              // AUTO only: fall back to the cooperative grid (same policy to rounding)
        if (int rc2 = pick_mode(IRLB200_MODE_AUTO, B, t->S, t->A, false, &mode)) return rc2;
    }
    if (mode == IRLB200_MODE_CTA) return launch_succ_cta<kOpBackward>(bt, B, (cudaStream_t)stream);
    return launch_succ_grid<kOpBackward>(bt.a, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int irlb200_soft_vi(const irlb200_tables *t, int B, const double *reward,
                               const double *phi, int phi_shared, double discount, double eps,
                               int max_sweeps, double *policy, double *value_out,
                               int32_t *n_iter, int32_t *status, int mode, void *stream) {
    if (int rc = check_tables(t, true, false)) return rc;
    if (B <= 0 || !reward || !phi || !policy) return fail(IRLB200_EINVAL, "soft_vi: bad argument");
    if (int rc = pick_mode(mode, B, t->S, t->A, false, &mode)) return rc;
    SuccBatch bt{};
    fill_succ(bt.a, t);
    bt.a.reward = reward; bt.a.phi = phi; bt.a.discount = discount; bt.a.eps = eps;
    bt.a.max_sweeps = max_sweeps; bt.a.policy = policy; bt.a.value = value_out;
    succ_strides(bt, t);
    bt.phi_stride = phi_shared ? 0 : (size_t)t->S;
    bt.n_iter = n_iter; bt.status = status; bt.out_stride = 1;
    if (mode == IRLB200_MODE_CTA) return launch_succ_cta<kOpSoftVI>(bt, B, (cudaStream_t)stream);
    return launch_succ_grid<kOpSoftVI>(bt.a, n_iter, status, (cudaStream_t)stream);
}

extern "C" int irlb200_value_iteration(const irlb200_tables *t, int B, const double *reward,
                                       double discount, double eps, int max_sweeps, int kind,
                                       double *value, int32_t *n_iter, int32_t *status,
                                       int mode, void *stream) {
    if (int rc = check_tables(t, true, false)) return rc;
    if (B <= 0 || !reward || !value) return fail(IRLB200_EINVAL, "value_iteration: bad argument");
    if (int rc = pick_mode(mode, B, t->S, t->A, false, &mode)) return rc;
    SuccBatch bt{};
    fill_succ(bt.a, t);
    bt.a.reward = reward; bt.a.discount = discount; bt.a.eps = eps; bt.a.max_sweeps = max_sweeps;
    bt.a.vi_mean = kind ? 1 : 0; bt.a.value = value;
    succ_strides(bt, t);
    bt.n_iter = n_iter; bt.status = status; bt.out_stride = 1;
    if (mode == IRLB200_MODE_CTA) return launch_succ_cta<kOpVI>(bt, B, (cudaStream_t)stream);
    return launch_succ_grid<kOpVI>(bt.a, n_iter, status, (cudaStream_t)stream);
}

static void fill_svf(SvfArgs &a, const irlb200_tables *t) {
    a = SvfArgs{};
    a.S = t->S; a.A = t->A; a.K = t->Kp;
    a.idx = t->pred_idx; a.p = t->pred_p;
    a.grid_n = t->stencil_n;
}

extern "C" int irlb200_svf(const irlb200_tables *t, int B, const double *p_initial, int p0_shared,
                           const uint8_t *terminal_mask, int mask_shared, const double *policy,
                           double eps, int max_sweeps, double *svf, const double *e_features,
                           int ef_shared, double *grad, int32_t *n_iter, int32_t *status, int mode,
                           void *stream) {
    if (int rc = check_tables(t, false, true)) return rc;
    if (B <= 0 || !p_initial || !terminal_mask || !policy || !svf) return fail(IRLB200_EINVAL, "svf: bad argument");
    if (grad && !e_features) return fail(IRLB200_EINVAL, "svf: grad requested without e_features");
    // thread-block-cluster mode: grid-stencil tables whose tile rows split evenly over <= 16 CTAs
    const int cl_size = (t->stencil_n > 0 && t->stencil_n * t->stencil_n == t->S && t->A == 4 && t->Kp == 5)
                            ? cluster_size_for(t->stencil_n) : 0;
    // AUTO: small batches of mid-size worlds (one world cannot fill an SM's latency budget), and every
    // world that does not fit one CTA's shared memory
    const bool want_cluster = mode == IRLB200_MODE_CLUSTER ||
                              (mode == IRLB200_MODE_AUTO && cl_size > 0 &&
                               ((t->S > 2048 && B <= 64) || t->S > 8192));
    if (mode == IRLB200_MODE_CLUSTER && cl_size == 0)
        return fail(IRLB200_ELIMIT, "cluster mode needs grid-stencil tables with n <= 128, n % 4 == 0");
    if (!want_cluster)
        if (int rc = pick_mode(mode, B, t->S, t->A, false, &mode)) return rc;
    SvfBatch bt{};
    fill_svf(bt.a, t);
    bt.a.p0 = p_initial; bt.a.term = terminal_mask; bt.a.policy = policy; bt.a.eps = eps;
    bt.a.max_sweeps = max_sweeps; bt.a.svf = svf; bt.a.e_features = e_features; bt.a.grad = grad;
    bt.tab_idx_stride = t->shared ? 0 : (size_t)t->Kp * t->S;
    bt.tab_p_stride = t->shared ? 0 : (size_t)t->A * t->Kp * t->S;
    bt.p0_stride = p0_shared ? 0 : (size_t)t->S;
    bt.term_stride = mask_shared ? 0 : (size_t)t->S;
    bt.ef_stride = ef_shared ? 0 : (size_t)t->S;
    bt.n_iter = n_iter; bt.status = status; bt.out_stride = 1;
    if (want_cluster) {
        if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
        int rc = launch_svf_cluster(bt, B, t->stencil_n, cl_size, (cudaStream_t)stream);
        if (rc != IRLB200_OK && cl_size > 8) {          // 16-CTA clusters are not schedulable everywhere
            cudaGetLastError();
            const int c8 = cluster_size_for(t->stencil_n, 8);
            if (c8 > 0) rc = launch_svf_cluster(bt, B, t->stencil_n, c8, (cudaStream_t)stream);
        }
        if (rc == IRLB200_OK || mode == IRLB200_MODE_CLUSTER) return rc;
        // AUTO only: this device cannot schedule the cluster shape -> cooperative grid (same results)
        cudaGetLastError();
        if (int rc2 = pick_mode(IRLB200_MODE_AUTO, B, t->S, t->A, false, &mode)) return rc2;
    }
    if (mode == IRLB200_MODE_CTA) return launch_svf_cta(bt, B, (cudaStream_t)stream);
    return launch_svf_grid(bt.a, n_iter, status, (cudaStream_t)stream);
}

extern "C" int irlb200_expected_svf(const irlb200_tables *t, int B, int causal,
                                    const double *reward, const double *p_initial, int p0_shared,
                                    const uint8_t *terminal_mask, const double *phi, int mask_shared,
                                    int n_backward, double discount, double eps_lap, double eps_svf,
                                    int max_sweeps, double *svf, const double *e_features, int ef_shared,
                                    double *grad, double *policy_out, int32_t *n_iter, int32_t *status,
                                    void *stream) {
    if (int rc = check_tables(t, true, true)) return rc;
    if (B <= 0 || !reward || !p_initial || !terminal_mask || !svf) return fail(IRLB200_EINVAL, "expected_svf: bad argument");
    if (causal && !phi) return fail(IRLB200_EINVAL, "expected_svf: causal needs phi");
    if (grad && !e_features) return fail(IRLB200_EINVAL, "expected_svf: grad requested without e_features");
    int mode = IRLB200_MODE_CTA;
    if (int rc = pick_mode(IRLB200_MODE_CTA, B, t->S, t->A, true, &mode)) return rc;
    StepBatch bt{};
    fill_succ(bt.s, t);
    bt.s.reward = reward; bt.s.phi = phi; bt.s.term = terminal_mask; bt.s.discount = discount;
    bt.s.eps = eps_lap; bt.s.n_sweeps = n_backward; bt.s.max_sweeps = max_sweeps;
    fill_svf(bt.f, t);
    bt.f.p0 = p_initial; bt.f.term = terminal_mask; bt.f.eps = eps_svf; bt.f.max_sweeps = max_sweeps;
    bt.f.svf = svf; bt.f.e_features = e_features; bt.f.grad = grad;
    bt.succ_idx_stride = t->shared ? 0 : (size_t)t->Ks * t->S;
    bt.succ_p_stride = t->shared ? 0 : (size_t)t->A * t->Ks * t->S;
    bt.pred_idx_stride = t->shared ? 0 : (size_t)t->Kp * t->S;
    bt.pred_p_stride = t->shared ? 0 : (size_t)t->A * t->Kp * t->S;
    bt.phi_stride = mask_shared ? 0 : (size_t)t->S;
    bt.term_stride = mask_shared ? 0 : (size_t)t->S;
    bt.p0_stride = p0_shared ? 0 : (size_t)t->S;
    bt.ef_stride = ef_shared ? 0 : (size_t)t->S;
    bt.n_iter = n_iter; bt.status = status; bt.policy_out = policy_out;
    return causal ? launch_step_cta<true>(bt, B, (cudaStream_t)stream)
                  : launch_step_cta<false>(bt, B, (cudaStream_t)stream);
}
