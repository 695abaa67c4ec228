// phases.cuh -- the four fixed-point loops of the hot path, written once
// against the Topo interface (topo.cuh).
//
//   succ_phase<OP>   sweeps over the SUCCESSOR table:
//       OpBackward   local_action_probabilities         maxent.py:119-159
//       OpSoftVI     local_causal_action_probabilities  maxent.py:279-341
//       OpVI         value_iteration                    solver.py:9-52 (+ :55-104)
//   svf_phase        sweeps over the PREDECESSOR table:
//                    expected_svf_from_policy           maxent.py:63-114
//
// Template parameters <A_T, K_T, SPT_T>:
//   SPT_T > 0   "register-resident": A_T, K_T are compile-time, every thread
//               owns SPT_T states (s = rank + k * nthreads) and keeps their
//               table rows in registers for the whole loop -- a sweep touches
//               no memory except the iterate itself.
//   SPT_T == 0  "streamed": every thread strides over the states and reads the
//               table rows from global memory each sweep (coalesced, slot-major).
//               A_T/K_T may be 0 => run-time A and K (any MDP).
// Both flavours evaluate exactly the same expression tree per state, so their
// results are bit-identical to each other.
#pragma once
#include "topo.cuh"

namespace irlb200 {

constexpr int kMaxDynA = 16;     // run-time A supported by the streamed flavour

enum OpKind : int { kOpBackward = 0, kOpSoftVI = 1, kOpVI = 2 };

struct SuccArgs {
    int S, A, K;                 // states, actions, successor slots
    const int32_t *idx;          // [K][S]
    const double *p;             // [A][K][S]
    const double *reward;        // [S]
    const double *phi;           // [S]       soft-VI terminal reward
    const uint8_t *term;         // [S]       backward: seeds zs
    double discount;
    double eps;
    int n_sweeps;                // backward: fixed sweep count (reference: 2*S)
    int max_sweeps;              // <= 0: unguarded
    int vi_mean;                 // VI: 1 = average over actions (solver.py:100)
    int grid_n;                  // > 0: successor offsets lie in {-n,-1,0,+1,+n} (grid stencil)
    int s_off;                   // streamed flavour: iterate index of local state 0 (slab mode; else 0).
                                 // Tables / per-state vectors are local (length S), neighbour indices and
                                 // the iterate are global.
    double *policy;              // [S][A] out (generic address: shared or global) or null
    double *policy2;             // optional second copy (global) or null
    double *value;               // [S] out (global) or null
};

struct SvfArgs {
    int S, A, K;                 // K = predecessor slots
    const int32_t *idx;          // [K][S]
    const double *p;             // [A][K][S]
    const double *p0;            // [S]
    const uint8_t *term;         // [S]
    const double *policy;        // [S][A] (generic address)
    double *w_scratch;           // [K][S] global scratch, streamed flavour only
    int grid_n;                  // > 0: predecessor offsets lie in {-n,-1,0,+1,+n} (grid stencil), n = grid_n
    int s_off;                   // streamed flavour: iterate index of local state 0 (slab mode; else 0);
                                 // policy / term are then indexed globally
    double eps;
    int max_sweeps;
    double *svf;                 // [S] out
    const double *e_features;    // [S] or null
    double *grad;                // [S] out or null: e_features - svf
};

// ---------------------------------------------------------------------------
// one state's update, shared by both flavours.
//   xv[j]  gathered iterate values, pr(a, j) table value accessor
// ---------------------------------------------------------------------------
template <int OP, int A_T, class PAcc, class XAcc>
__device__ __forceinline__ double succ_update(int A, int K, PAcc pr, XAcc xv, double c0, double c1,
                                              double discount, int vi_mean, double *q_out) {
    constexpr int QN = A_T > 0 ? A_T : kMaxDynA;
    double q[QN];
    const int An = A_T > 0 ? A_T : A;
#pragma unroll
    for (int a = 0; a < An; ++a) {
        double dot = 0.0;
        for (int j = 0; j < K; ++j) dot = fma(pr(a, j), xv(j), dot);     // P_a.dot(x), ascending s'
        if (OP == kOpBackward) q[a] = c0 * dot;                          // er * P_a.dot(zs)      :155
        else if (OP == kOpSoftVI) q[a] = c0 + discount * dot;            // r + g * P_a.dot(v)    :329
        else q[a] = discount * dot;                                      // g * (P_a @ v)   solver:44
    }
    double x;
    if (OP == kOpBackward) {
        x = q[0];
#pragma unroll
        for (int a = 1; a < An; ++a) x += q[a];                          // za.sum(axis=1)        :156
    } else if (OP == kOpSoftVI) {
        // v = softmax(...softmax(softmax(phi, q_0), q_1)..., q_{A-1})  (maxent.py:331-333) is
        // log(e^phi + sum_a e^{q_a}).  The reference folds it pairwise: A dependent exp+log pairs,
        // ~2 500 cycles of FP64 latency per sweep.  Evaluated here as m + log(sum_i exp(x_i - m)) with
        // m = max_i x_i: the exps are independent and there is one log (agrees with the fold to a few
        // ulp; the tests hold policies to 1e-10 and sweep counts exactly).  A non-finite maximum
        // (+inf, NaN, or everything -inf) takes the reference's fold verbatim.
        double m = c1;
#pragma unroll
        for (int a = 0; a < An; ++a) m = max_nan(m, q[a]);
        if (fabs(m) < INFINITY) {
            double ssum = 0.0;
            if (c1 != -INFINITY) ssum = exp(c1 - m);                     // phi = -inf off the terminals
#pragma unroll
            for (int a = 0; a < An; ++a) ssum += exp(q[a] - m);
            x = m + log(ssum);
        } else {
            x = c1;                                                      // v = reward_terminal   :331
#pragma unroll 1
            for (int a = 0; a < An; ++a) x = softmax2(x, q[a]);          //                       :332-333
        }
    } else {
        if (vi_mean) {
            x = q[0];
#pragma unroll
            for (int a = 1; a < An; ++a) x += q[a];
            x = c0 + x / (double)An;                                     // solver.py:100
        } else {
            x = q[0];
#pragma unroll
            for (int a = 1; a < An; ++a) x = max_nan(x, q[a]);
            x = c0 + x;                                                  // solver.py:47
        }
    }
    if (q_out) {
#pragma unroll
        for (int a = 0; a < An; ++a) q_out[a] = q[a];
    }
    return x;
}

template <int OP>
__device__ __forceinline__ double succ_init_value(bool is_terminal) {
    if (OP == kOpBackward) return is_terminal ? 1.0 : 0.0;               // maxent.py:146-147
    if (OP == kOpSoftVI) return kNegHuge;                                // maxent.py:323
    return 0.0;                                                          // solver.py:29
}

// rescale period of the backward pass: the largest entry changes per sweep by
// at most a factor A * e^{max r} (and by at least e^{min r} * p), so R sweeps
// move the exponent by < ~256 and the values stay far from the ends of the
// double range.  The scaling itself is by an exact power of two.
__device__ __forceinline__ int backward_rescale_period(double max_abs_r, int A) {
    if (!(max_abs_r < 1e300)) return 1;
    double bits = log2((double)A) + max_abs_r * 1.4426950408889634 + 8.0;
    double R = floor(256.0 / bits);
    return R < 1.0 ? 1 : (R > 64.0 ? 64 : (int)R);
}

// ---------------------------------------------------------------------------
// successor-table phase
// ---------------------------------------------------------------------------
template <class Topo, int OP, int A_T, int K_T, int SPT_T>
__device__ void succ_phase(Topo &tp, const SuccArgs &a, int *n_iter_out, int *status_out) {
    const int S = a.S;
    const int A = A_T > 0 ? A_T : a.A;
    const int K = K_T > 0 ? K_T : a.K;
    const int NT = tp.nthreads();
    const int r0 = tp.rank();
    constexpr bool REG = SPT_T > 0;
    constexpr int SPT = REG ? SPT_T : 1;
    constexpr int AR = REG ? A_T : 1, KR = REG ? K_T : 1;

    // register-resident rows (REG flavour only)
    double pr[SPT][AR][KR];
    int ix[SPT][KR];
    double c0[SPT], c1[SPT], cur[SPT];

    tp.begin_phase();

    // ---- prologue: constants, initial iterate --------------------------------
    double max_abs_r = 0.0;
    if (REG) {
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int s = r0 + k * NT;
            const bool act = s < S;
#pragma unroll
            for (int j = 0; j < KR; ++j) {
                ix[k][j] = act ? a.idx[(size_t)j * S + s] : 0;
#pragma unroll
                for (int aa = 0; aa < AR; ++aa)
                    pr[k][aa][j] = act ? a.p[((size_t)aa * K + j) * S + s] : 0.0;
            }
            const double r = act ? a.reward[s] : 0.0;
            max_abs_r = fmax(max_abs_r, fabs(r));
            c0[k] = (OP == kOpBackward) ? exp(r) : r;                    // er = np.exp(reward)   :142
            c1[k] = (OP == kOpSoftVI && act) ? a.phi[s] : 0.0;
            cur[k] = succ_init_value<OP>(OP == kOpBackward && act && a.term[s]);
            if (act) tp.store(0, s, cur[k]);
        }
    } else {
        for (int s = r0; s < S; s += NT) {
            max_abs_r = fmax(max_abs_r, fabs(a.reward[s]));
            tp.store(0, a.s_off + s, succ_init_value<OP>(OP == kOpBackward && a.term[s]));
        }
    }
    int R = 0;
    if (OP == kOpBackward) R = backward_rescale_period(tp.reduce_max(max_abs_r), A);
    tp.sync();

    // ---- sweeps --------------------------------------------------------------
    int n = 0, status = IRLB200_ST_CONVERGED;
    const bool fixed = (OP == kOpBackward);
    const int limit = fixed ? a.n_sweeps : a.max_sweeps;
    if (!fixed || a.n_sweeps > 0) {
        for (;;) {
            const int b = n & 1;
            Vote v;
            v.reset();
            double local_max = 0.0;
            if (REG) {
#pragma unroll
                for (int k = 0; k < SPT; ++k) {
                    const int s = r0 + k * NT;
                    if (s < S) {
                        double xv[KR];
#pragma unroll
                        for (int j = 0; j < KR; ++j) xv[j] = tp.load(b, ix[k][j]);
                        const double x = succ_update<OP, A_T>(
                            A, K, [&](int aa, int j) { return pr[k][aa][j]; },
                            [&](int j) { return xv[j]; }, c0[k], c1[k], a.discount, a.vi_mean, nullptr);
                        if (!fixed) v.add(x, cur[k], a.eps);
                        cur[k] = x;
                        local_max = fmax(local_max, x);
                        tp.store(b ^ 1, s, x);
                    }
                }
            } else {
                for (int s = r0; s < S; s += NT) {
                    const double r = a.reward[s];
                    const double k0 = (OP == kOpBackward) ? exp(r) : r;
                    const double k1 = (OP == kOpSoftVI) ? a.phi[s] : 0.0;
                    const double x = succ_update<OP, A_T>(
                        A, K, [&](int aa, int j) { return __ldg(a.p + ((size_t)aa * K + j) * S + s); },
                        [&](int j) { return tp.load(b, __ldg(a.idx + (size_t)j * S + s)); }, k0, k1,
                        a.discount, a.vi_mean, nullptr);
                    if (!fixed) v.add(x, tp.load(b, a.s_off + s), a.eps);
                    local_max = fmax(local_max, x);
                    tp.store(b ^ 1, a.s_off + s, x);
                }
            }
            ++n;
            if (fixed) {
                if (n >= limit) { tp.sync(); break; }
                if (n % R == 0) {
                    // exact power-of-two rescale of the new iterate (range extension)
                    const double m = tp.reduce_max(local_max);
                    if (m > 0.0 && m < INFINITY) {
                        const int e = frexp_exponent(m);
                        if (REG) {
#pragma unroll
                            for (int k = 0; k < SPT; ++k) {
                                const int s = r0 + k * NT;
                                if (s < S) { cur[k] = ldexp(cur[k], -e); tp.store(b ^ 1, s, cur[k]); }
                            }
                        } else {
                            for (int s = r0; s < S; s += NT)
                                tp.store(b ^ 1, a.s_off + s, ldexp(tp.load(b ^ 1, a.s_off + s), -e));
                        }
                    }
                }
                tp.sync();
            } else {
                const int st = tp.vote(v);
                if (st != kContinue) { status = st; break; }
                if (limit > 0 && n >= limit) { status = IRLB200_ST_MAXSWEEPS; break; }
            }
        }
    }

    // ---- epilogue: outputs from the last sweep --------------------------------
    // buffer (n-1)&1 still holds the iterate the last sweep read, buffer n&1 the
    // one it wrote; recomputing the last sweep's per-action terms from the former
    // reproduces them bit for bit.
    if (n > 0 && (a.policy || a.policy2) && OP != kOpVI) {
        const int b = (n - 1) & 1;
        auto emit = [&](int s, const double *q, double x) {
            for (int aa = 0; aa < A; ++aa) {
                const double pol = (OP == kOpBackward) ? q[aa] / x          // za / zs[:, None]  :159
                                                       : exp(q[aa] - x);    // exp(q - v[:,None]):341
                if (a.policy) a.policy[(size_t)s * A + aa] = pol;
                if (a.policy2) a.policy2[(size_t)s * A + aa] = pol;
            }
        };
        constexpr int QN = A_T > 0 ? A_T : kMaxDynA;
        if (REG) {
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
                const int s = r0 + k * NT;
                if (s < S) {
                    double xv[KR], q[QN];
#pragma unroll
                    for (int j = 0; j < KR; ++j) xv[j] = tp.load(b, ix[k][j]);
                    const double x = succ_update<OP, A_T>(
                        A, K, [&](int aa, int j) { return pr[k][aa][j]; },
                        [&](int j) { return xv[j]; }, c0[k], c1[k], a.discount, a.vi_mean, q);
                    emit(s, q, x);
                }
            }
        } else {
            for (int s = r0; s < S; s += NT) {
                double q[QN];
                const double r = a.reward[s];
                const double k0 = (OP == kOpBackward) ? exp(r) : r;
                const double k1 = (OP == kOpSoftVI) ? a.phi[s] : 0.0;
                const double x = succ_update<OP, A_T>(
                    A, K, [&](int aa, int j) { return __ldg(a.p + ((size_t)aa * K + j) * S + s); },
                    [&](int j) { return tp.load(b, __ldg(a.idx + (size_t)j * S + s)); }, k0, k1,
                    a.discount, a.vi_mean, q);
                emit(s, q, x);
            }
        }
    }
    if (a.value) {
        const int b = n & 1;
        for (int s = r0; s < S; s += NT) a.value[s] = tp.load(b, (REG ? 0 : a.s_off) + s);
    }
    if (r0 == 0) {
        if (n_iter_out) *n_iter_out = n;
        if (status_out) *status_out = status;
    }
    tp.sync();
}

// ---------------------------------------------------------------------------
// predecessor-table phase (forward state-visitation pass)
//   d'[s'] = p0[s'] + sum_j W[j][s'] * d[pred_j(s')]
//   W[j][s'] = sum_a P[pred_j, s', a] * policy[pred_j, a], 0 if pred_j is terminal
//   (maxent.py:98-99 drops the outgoing rows of terminal states; :109-110)
// ---------------------------------------------------------------------------
template <class Topo, int A_T, int K_T, int SPT_T>
__device__ void svf_phase(Topo &tp, const SvfArgs &a, int *n_iter_out, int *status_out) {
    const int S = a.S;
    const int A = A_T > 0 ? A_T : a.A;
    const int K = K_T > 0 ? K_T : a.K;
    const int NT = tp.nthreads();
    const int r0 = tp.rank();
    constexpr bool REG = SPT_T > 0;
    constexpr int SPT = REG ? SPT_T : 1;
    constexpr int KR = REG ? K_T : 1;

    double w[SPT][KR];
    int ix[SPT][KR];
    double p0r[SPT], cur[SPT];

    tp.begin_phase();

    auto weight = [&](int s, int j, int pred) -> double {
        double acc = 0.0;
        for (int aa = 0; aa < A; ++aa)
            acc = fma(__ldg(a.p + ((size_t)aa * K + j) * S + s), a.policy[(size_t)pred * A + aa], acc);
        return a.term[pred] ? 0.0 : acc;
    };

    if (REG) {
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int s = r0 + k * NT;
            const bool act = s < S;
#pragma unroll
            for (int j = 0; j < KR; ++j) {
                ix[k][j] = act ? a.idx[(size_t)j * S + s] : 0;
                w[k][j] = act ? weight(s, j, ix[k][j]) : 0.0;
            }
            p0r[k] = act ? a.p0[s] : 0.0;
            cur[k] = 0.0;
            if (act) tp.store(0, s, 0.0);                               // d = np.zeros       :105
        }
    } else {
        for (int s = r0; s < S; s += NT) {
            for (int j = 0; j < K; ++j)
                a.w_scratch[(size_t)j * S + s] = weight(s, j, a.idx[(size_t)j * S + s]);
            tp.store(0, a.s_off + s, 0.0);
        }
    }
    tp.sync();

    int n = 0, status = IRLB200_ST_CONVERGED;
    for (;;) {
        const int b = n & 1;
        Vote v;
        v.reset();
        if (REG) {
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
                const int s = r0 + k * NT;
                if (s < S) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < KR; ++j) acc = fma(w[k][j], tp.load(b, ix[k][j]), acc);
                    const double x = p0r[k] + acc;                      // p_initial + sum     :110
                    v.add(x, cur[k], a.eps);
                    cur[k] = x;
                    tp.store(b ^ 1, s, x);
                }
            }
        } else {
            for (int s = r0; s < S; s += NT) {
                double acc = 0.0;
                for (int j = 0; j < K; ++j)
                    acc = fma(a.w_scratch[(size_t)j * S + s], tp.load(b, __ldg(a.idx + (size_t)j * S + s)), acc);
                const double x = __ldg(a.p0 + s) + acc;
                v.add(x, tp.load(b, a.s_off + s), a.eps);
                tp.store(b ^ 1, a.s_off + s, x);
            }
        }
        ++n;
        const int st = tp.vote(v);
        if (st != kContinue) { status = st; break; }
        if (a.max_sweeps > 0 && n >= a.max_sweeps) { status = IRLB200_ST_MAXSWEEPS; break; }
    }

    const int b = n & 1;
    for (int s = r0; s < S; s += NT) {
        const double d = tp.load(b, (REG ? 0 : a.s_off) + s);
        a.svf[s] = d;
        if (a.grad) a.grad[s] = a.e_features[s] - d;                    // maxent.py:248, features = I
    }
    if (r0 == 0) {
        if (n_iter_out) *n_iter_out = n;
        if (status_out) *status_out = status;
    }
    tp.sync();
}

}  // namespace irlb200
