// kernels_tiled.cuh -- stencil-tiled kernels for grid worlds, one CTA per world: forward pass and merged-weight backward pass.
#pragma once
#include "batch_args.cuh"

namespace irlb200 {

// ---------------------------------------------------------------------------
// Stencil-tiled forward pass for grid worlds (predecessor offsets within {-n,-1,0,+1,+n}).
//
// The generic gather (svf_cta_fast_kernel, kernels_cta.cuh) is bound by the shared-memory datapath: 5 LDS.64 per state =
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
// ---------------------------------------------------------------------------
// Per-cell prologues shared by every stencil kernel (one CTA, cluster, push): the table row of a cell
// folded into five weights by stencil position (s-n, s-1, s, s+1, s+n).  At most one table entry per
// position carries weight; padding entries add +0.
// ---------------------------------------------------------------------------
// forward pass: W[j] = sum_a P[pred_j, s, a] * policy[pred_j, a], 0 when pred_j is terminal (maxent.py:98-110)
template <int A, int K>
__device__ __forceinline__ void svf_stencil_weights(const SvfArgs &a, const int S, const int s, const int n,
                                                    const bool live, double (&w)[5]) {
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = 0.0;
    if (!live) return;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int pred = a.idx[(size_t)j * S + s];
        double acc = 0.0;
#pragma unroll
        for (int aa = 0; aa < A; ++aa)
            acc = fma(__ldg(a.p + ((size_t)aa * K + j) * S + s), a.policy[(size_t)pred * A + aa], acc);
        if (a.term[pred]) acc = 0.0;
        const int off = pred - s;
        w[0] += (off == -n) ? acc : 0.0;
        w[1] += (off == -1) ? acc : 0.0;
        w[2] += (off == 0) ? acc : 0.0;
        w[3] += (off == 1) ? acc : 0.0;
        w[4] += (off == n) ? acc : 0.0;
    }
}

// backward pass: merged weights er * sum_a P[s, succ_j, a]; returns the start value zs0 of the cell and
// folds |reward| into max_abs_r (maxent.py:142, :146-147, :155-156)
template <int A, int K>
__device__ __forceinline__ double backward_stencil_weights(const SuccArgs &a, const int S, const int s, const int n,
                                                           const bool live, double (&w)[5], double &max_abs_r) {
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = 0.0;
    if (!live) return 0.0;
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
        w[0] += (off == -n) ? q : 0.0;
        w[1] += (off == -1) ? q : 0.0;
        w[2] += (off == 0) ? q : 0.0;
        w[3] += (off == 1) ? q : 0.0;
        w[4] += (off == n) ? q : 0.0;
    }
    return a.term[s] ? 1.0 : 0.0;                               // zs[terminal] = 1.0  :146-147
}

// The arithmetic of one sweep over a thread's TY x TX tile: per cell the ascending-neighbour FMA chain
// (s-n, s-1, s, s+1, s+n) of the ELL kernels, neighbours from the tile itself or from the halo values
// `up` / `dn` / `lf` / `rt`; ADD_P0: `p_initial + sum` (forward pass, maxent.py:110), else the plain
// sum (merged-weight backward sweep, maxent.py:155-156).  Stated once so that every stencil kernel is
// bitwise identical to the others.
template <int TY, int TX, bool ADD_P0>
__device__ __forceinline__ void stencil_tile_update(const double (&w)[TY * TX][5], const double (&p0r)[TY * TX],
                                                    const double (&cur)[TY * TX], const double (&up)[TX],
                                                    const double (&dn)[TX], const double (&lf)[TY],
                                                    const double (&rt)[TY], double (&x)[TY * TX]) {
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
            x[c] = ADD_P0 ? p0r[c] + acc : acc;
        }
}

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
    stencil_tile_update<TY, TX, true>(w, p0r, cur, up, dn, lf, rt, x);
#pragma unroll
    for (int c = 0; c < TY * TX; ++c) *reinterpret_cast<double *>(smem + own + 8 * c + OFF_W) = x[c];
}

// One sweep src -> dst plus the stop rule.  Returns kContinue or the final status.
template <int TY, int TX, int MAXT, int OFF_R, int OFF_W>
__device__ __forceinline__ int svf_grid5_step(unsigned char *smem, uint32_t own, uint32_t nb_up, uint32_t nb_dn,
                                              uint32_t nb_lf, uint32_t nb_rt, const double (&w)[TY * TX][5],
                                              const double (&p0r)[TY * TX], const double (&src)[TY * TX],
                                              double (&dst)[TY * TX], const double eps, const int limit, int &nsw,
                                              int *flag) {
    constexpr int C = TY * TX;
    svf_grid5_sweep<TY, TX, MAXT, OFF_R, OFF_W>(smem, own, nb_up, nb_dn, nb_lf, nb_rt, w, p0r, src, dst);
    ++nsw;
    if (!__syncthreads_or(!(fabs(dst[0] - src[0]) <= eps) ? 1 : 0)) {
        bool go = false;
#pragma unroll
        for (int c = 1; c < C; ++c) go |= !(fabs(dst[c] - src[c]) <= eps);   // |diff| > eps, or NaN
        if (!__syncthreads_or(go ? 1 : 0)) return IRLB200_ST_CONVERGED;     // delta <= eps
    }
    if ((nsw & 15) == 0) {
        bool bad = false;
#pragma unroll
        for (int c = 0; c < C; ++c) bad |= (dst[c] - dst[c]) != 0.0;
        if (bad) *flag = 1;
        __syncthreads();
        if (*flag) return IRLB200_ST_NONFINITE;
    }
    if (nsw >= limit) return IRLB200_ST_MAXSWEEPS;
    return kContinue;
}

template <int TY, int TX, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) svf_grid5_kernel(const SvfBatch bt, const int n) {
    using Cfg = Grid5Cfg<TY, TX, MAXT>;
    constexpr int C = Cfg::C, K = 5, A = 4, STRIDE = Cfg::STRIDE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *flag = reinterpret_cast<int *>(smem_raw + 2 * STRIDE);

    SvfArgs a = bt.a;
    const size_t prob = svf_problem(bt);
    offset_svf(a, bt, prob);
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
            svf_stencil_weights<A, K>(a, S, s, n, live, w[c]);
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
    // Two sweeps per trip, ping-ponging between two register arrays (cur -> x, x -> cur): no copies and no
    // parity branch in the loop (they were 20 of the ~100 instructions of a thread-sweep).
    int nsw = 0, status;
    double x[C];
    for (;;) {
        status = svf_grid5_step<TY, TX, MAXT, 0, STRIDE>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, p0r, cur, x, eps, limit, nsw, flag);
        if (status != kContinue) {
#pragma unroll
            for (int c = 0; c < C; ++c) cur[c] = x[c];
            break;
        }
        status = svf_grid5_step<TY, TX, MAXT, STRIDE, 0>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, p0r, x, cur, eps, limit, nsw, flag);
        if (status != kContinue) break;
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
        if (bt.n_iter) bt.n_iter[prob * bt.out_stride] = nsw;
        if (bt.status) bt.status[prob * bt.out_stride] = status;
    }
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
template <int TY, int TX, int MAXT, int OFF_R, int OFF_W, bool WANT_MAX>
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
    stencil_tile_update<TY, TX, false>(w, cur, cur, up, dn, lf, rt, x);
#pragma unroll
    for (int c = 0; c < TY * TX; ++c) {
        *reinterpret_cast<double *>(smem + own + 8 * c + OFF_W) = x[c];
        cur[c] = x[c];
        if (WANT_MAX) m = fmax(m, x[c]);                  // only the rescale sweeps need the maximum
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
            const double z0 = backward_stencil_weights<A, K>(a, S, s, n, live, w[c], max_abs_r);
            cur[c] = z0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c) = z0;
            *reinterpret_cast<double *>(smem_raw + own + 8 * c + STRIDE) = 0.0;
        }
    const int R = backward_rescale_period(block_max(max_abs_r, scratch), A);
    __syncthreads();

    // ---- n_sweeps - 1 merged-weight sweeps -------------------------------------------------
    const int n_lin = a.n_sweeps - 1;
    int until_rescale = R;                                  // countdown instead of (t + 1) % R
    for (int t = 0; t < n_lin; ++t) {
        const bool rescale = --until_rescale == 0 && t + 1 < n_lin;
        if (until_rescale == 0) until_rescale = R;
        double m = 0.0;
        if (rescale) {
            m = (t & 1)
                ? lin_grid5_sweep<TY, TX, MAXT, STRIDE, 0, true>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, cur)
                : lin_grid5_sweep<TY, TX, MAXT, 0, STRIDE, true>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, cur);
        } else if (t & 1) {
            lin_grid5_sweep<TY, TX, MAXT, STRIDE, 0, false>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, cur);
        } else {
            lin_grid5_sweep<TY, TX, MAXT, 0, STRIDE, false>(smem_raw, own, nb_up, nb_dn, nb_lf, nb_rt, w, cur);
        }
        if (rescale) {
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


}  // namespace irlb200
