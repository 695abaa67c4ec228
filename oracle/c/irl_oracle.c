/*
 * irl_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the reference's hot path over per-(state, action) successor lists
 * (ELL with K slots: sidx[(s*A + a)*K + k], sp[...], unused slots have probability 0).  Each loop
 * follows the reference line by line with `P_a.dot(x)` / `P_a.T.dot(x)` written out over the
 * non-zeros of P_a in ascending state order:
 *
 *   oracle_backward   local_action_probabilities          /root/reference/src/maxent.py:119-159
 *   oracle_soft_vi    local_causal_action_probabilities   maxent.py:279-341 (softmax :260-276)
 *   oracle_svf        expected_svf_from_policy            maxent.py:63-114
 *   oracle_vi         value_iteration                     solver.py:9-52
 *   oracle_batch_maxent_step   B independent worlds: backward + forward pass (maxent.py:162-193),
 *                     one world per OpenMP thread -- the "optimised CPU" line of bench.py.
 *
 * Never linked into, loaded by or called from the product (irl-maxent_b200/).  Pinned through
 * tests/test_oracle_c.py against the numpy restatement, which is pinned bit-for-bit to the
 * unmodified reference (tests/golden).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEG_HUGE (-1e200)   /* maxent.py:323 */

static double softmax2(double x1, double x2) {            /* maxent.py:274-276 */
    double hi = x1 > x2 ? x1 : x2, lo = x1 > x2 ? x2 : x1;
    if (x1 != x1 || x2 != x2) return x1 + x2;             /* np.maximum propagates NaN */
    return hi + log(1.0 + exp(lo - hi));
}

/* maxent.py:119-159.  rescale != 0: exact power-of-two rescale per sweep (range extension). */
int oracle_backward(int S, int A, int K, const int *sidx, const double *sp, const unsigned char *term,
                    const double *reward, long n_sweeps, int rescale, double *policy) {
    double *er = malloc(sizeof(double) * S), *zs = calloc(S, sizeof(double)), *zn = malloc(sizeof(double) * S);
    double *za = malloc(sizeof(double) * (size_t)S * A);
    if (!er || !zs || !zn || !za) return -1;
    for (int s = 0; s < S; ++s) { er[s] = exp(reward[s]); zs[s] = term[s] ? 1.0 : 0.0; }     /* :142,:146-147 */
    for (long t = 0; t < n_sweeps; ++t) {                                                       /* :154 */
        double m = 0.0;
#pragma omp parallel for schedule(static) reduction(max : m) if (S >= 4096)
        for (int s = 0; s < S; ++s) {
            double sum = 0.0;
            for (int a = 0; a < A; ++a) {
                double dot = 0.0;
                const size_t o = ((size_t)s * A + a) * K;
                for (int k = 0; k < K; ++k) dot += sp[o + k] * zs[sidx[o + k]];
                za[(size_t)s * A + a] = er[s] * dot;                                            /* :155 */
                sum = a == 0 ? za[(size_t)s * A] : sum + za[(size_t)s * A + a];                 /* :156 */
            }
            zn[s] = sum;
            if (sum > m) m = sum;
        }
        if (rescale && m > 0.0 && m < INFINITY) {
            int e; frexp(m, &e);
            for (int s = 0; s < S; ++s) {
                zn[s] = ldexp(zn[s], -e);
                for (int a = 0; a < A; ++a) za[(size_t)s * A + a] = ldexp(za[(size_t)s * A + a], -e);
            }
        }
        double *tmp = zs; zs = zn; zn = tmp;
    }
    for (int s = 0; s < S; ++s)
        for (int a = 0; a < A; ++a) policy[(size_t)s * A + a] = za[(size_t)s * A + a] / zs[s];  /* :159 */
    free(er); free(zs); free(zn); free(za);
    return 0;
}

/* maxent.py:279-341.  phi: terminal reward function (0 at terminals, -inf elsewhere, or the caller's). */
int oracle_soft_vi(int S, int A, int K, const int *sidx, const double *sp, const double *phi,
                   const double *reward, double discount, double eps, long max_sweeps,
                   double *policy, double *value, long *n_out) {
    double *v = malloc(sizeof(double) * S), *vn = malloc(sizeof(double) * S), *q = malloc(sizeof(double) * (size_t)S * A);
    if (!v || !vn || !q) return -1;
    for (int s = 0; s < S; ++s) v[s] = NEG_HUGE;                                                /* :323 */
    long n = 0;
    double delta = INFINITY;
    while (delta > eps) {                                                                       /* :326 */
        delta = 0.0;
        int nan = 0;
        for (int s = 0; s < S; ++s) {
            double x = phi[s];                                                                  /* :331 */
            for (int a = 0; a < A; ++a) {
                double dot = 0.0;
                const size_t o = ((size_t)s * A + a) * K;
                for (int k = 0; k < K; ++k) dot += sp[o + k] * v[sidx[o + k]];
                q[(size_t)s * A + a] = reward[s] + discount * dot;                              /* :329 */
                x = softmax2(x, q[(size_t)s * A + a]);                                          /* :332-333 */
            }
            vn[s] = x;
            double diff = fabs(x - v[s]);                                                       /* :338 */
            if (diff != diff) nan = 1; else if (diff > delta) delta = diff;
        }
        if (nan) delta = NAN;
        double *tmp = v; v = vn; vn = tmp;
        ++n;
        if (max_sweeps > 0 && n >= max_sweeps) break;
    }
    for (int s = 0; s < S; ++s) {
        if (value) value[s] = v[s];
        for (int a = 0; a < A; ++a) policy[(size_t)s * A + a] = exp(q[(size_t)s * A + a] - v[s]);  /* :341 */
    }
    if (n_out) *n_out = n;
    free(v); free(vn); free(q);
    return 0;
}

/* maxent.py:63-114: per action the transposed product P_a.T.dot(policy[:, a] * d) (terminal rows of P
 * zeroed, :98-99), summed over the actions, plus p_initial (:109-110). */
int oracle_svf(int S, int A, int K, const int *sidx, const double *sp, const double *p0,
               const unsigned char *term, const double *policy, double eps, long max_sweeps,
               double *d_out, long *n_out) {
    double *d = calloc(S, sizeof(double)), *part = malloc(sizeof(double) * (size_t)S * A), *dn = malloc(sizeof(double) * S);
    if (!d || !part || !dn) return -1;
    long n = 0;
    double delta = INFINITY;
    while (delta > eps) {                                                                       /* :108 */
        memset(part, 0, sizeof(double) * (size_t)S * A);
        for (int a = 0; a < A; ++a) {
            double *pa = part + (size_t)a * S;
            for (int s = 0; s < S; ++s) {                      /* ascending s: the order of a row-major dgemv^T */
                if (term[s]) continue;                                                          /* :99 */
                const double m = policy[(size_t)s * A + a] * d[s];                              /* :109 */
                const size_t o = ((size_t)s * A + a) * K;
                for (int k = 0; k < K; ++k)
                    if (sp[o + k] != 0.0) pa[sidx[o + k]] += sp[o + k] * m;
            }
        }
        delta = 0.0;
        int nan = 0;
        for (int s = 0; s < S; ++s) {
            double sum = part[s];
            for (int a = 1; a < A; ++a) sum += part[(size_t)a * S + s];                         /* .sum(axis=0) */
            const double x = p0[s] + sum;                                                       /* :110 */
            const double diff = fabs(x - d[s]);                                                 /* :112 */
            if (diff != diff) nan = 1; else if (diff > delta) delta = diff;
            dn[s] = x;
        }
        if (nan) delta = NAN;
        double *tmp = d; d = dn; dn = tmp;
        ++n;
        if (max_sweeps > 0 && n >= max_sweeps) break;
    }
    memcpy(d_out, d, sizeof(double) * S);
    if (n_out) *n_out = n;
    free(d); free(part); free(dn);
    return 0;
}

/* The same loop as oracle_svf with the states of a sweep spread over OpenMP threads (full-size C3 checks:
 * 2.4e5 sweeps of 16 384 states).  The scatter `pa[s'] += sp * m` above visits the sources in ascending s;
 * here every target s' gathers its own contributions from a predecessor list built in that same order,
 * so each output is the same chain of additions and the result is BITWISE that of oracle_svf
 * (tests/test_oracle_c.py holds them equal). */
int oracle_svf_mt(int S, int A, int K, const int *sidx, const double *sp, const double *p0,
                  const unsigned char *term, const double *policy, double eps, long max_sweeps,
                  double *d_out, long *n_out) {
    /* predecessor lists per (action, target): sources ascending, terminal sources and zero entries dropped (:99) */
    size_t *off = calloc((size_t)S * A + 1, sizeof(size_t));
    if (!off) return -1;
    for (int s = 0; s < S; ++s) {
        if (term[s]) continue;
        for (int a = 0; a < A; ++a) {
            const size_t o = ((size_t)s * A + a) * K;
            for (int k = 0; k < K; ++k)
                if (sp[o + k] != 0.0) ++off[(size_t)a * S + sidx[o + k] + 1];
        }
    }
    for (size_t i = 0; i < (size_t)S * A; ++i) off[i + 1] += off[i];
    const size_t nnz = off[(size_t)S * A];
    int *src = malloc(sizeof(int) * (nnz ? nnz : 1));
    double *prob = malloc(sizeof(double) * (nnz ? nnz : 1));
    size_t *fill = malloc(sizeof(size_t) * (size_t)S * A);
    double *d = calloc(S, sizeof(double)), *dn = malloc(sizeof(double) * S);
    if (!src || !prob || !fill || !d || !dn) return -1;
    memcpy(fill, off, sizeof(size_t) * (size_t)S * A);
    for (int s = 0; s < S; ++s) {                              /* ascending s, as the scatter visits them */
        if (term[s]) continue;
        for (int a = 0; a < A; ++a) {
            const size_t o = ((size_t)s * A + a) * K;
            for (int k = 0; k < K; ++k)
                if (sp[o + k] != 0.0) {
                    const size_t at = fill[(size_t)a * S + sidx[o + k]]++;
                    src[at] = s;
                    prob[at] = sp[o + k];
                }
        }
    }
    long n = 0;
    int done = 0, nanflag = 0;
    double dmax = 0.0;
#pragma omp parallel
    {
        while (!done) {                                                                         /* :108 */
#pragma omp single
            { dmax = 0.0; nanflag = 0; }
#pragma omp for schedule(static) reduction(max : dmax) reduction(| : nanflag)
            for (int t = 0; t < S; ++t) {
                double sum = 0.0;
                for (int a = 0; a < A; ++a) {
                    double pa = 0.0;
                    for (size_t e = off[(size_t)a * S + t]; e < off[(size_t)a * S + t + 1]; ++e) {
                        const double m = policy[(size_t)src[e] * A + a] * d[src[e]];            /* :109 */
                        pa += prob[e] * m;
                    }
                    sum = a == 0 ? pa : sum + pa;                                               /* .sum(axis=0) */
                }
                const double x = p0[t] + sum;                                                   /* :110 */
                const double diff = fabs(x - d[t]);                                             /* :112 */
                if (diff != diff) nanflag |= 1; else if (diff > dmax) dmax = diff;
                dn[t] = x;
            }
#pragma omp single
            {
                const double delta = nanflag ? NAN : dmax;
                double *tmp = d; d = dn; dn = tmp;
                ++n;
                done = !(delta > eps) || (max_sweeps > 0 && n >= max_sweeps);
            }
        }
    }
    memcpy(d_out, d, sizeof(double) * S);
    if (n_out) *n_out = n;
    free(off); free(src); free(prob); free(fill); free(d); free(dn);
    return 0;
}

/* solver.py:9-52 */
int oracle_vi(int S, int A, int K, const int *sidx, const double *sp, const double *reward, double discount,
              double eps, long max_sweeps, double *v_out, long *n_out) {
    double *v = calloc(S, sizeof(double)), *vn = malloc(sizeof(double) * S);
    if (!v || !vn) return -1;
    long n = 0;
    double delta = INFINITY;
    while (delta > eps) {                                                                       /* :40 */
        delta = 0.0;
        for (int s = 0; s < S; ++s) {
            double best = 0.0;
            for (int a = 0; a < A; ++a) {
                double dot = 0.0;
                const size_t o = ((size_t)s * A + a) * K;
                for (int k = 0; k < K; ++k) dot += sp[o + k] * v[sidx[o + k]];
                const double qa = discount * dot;                                               /* :44 */
                if (a == 0 || qa > best) best = qa;                                             /* :47 */
            }
            vn[s] = reward[s] + best;
            const double diff = fabs(v[s] - vn[s]);                                             /* :50 */
            if (diff > delta) delta = diff;
        }
        double *tmp = v; v = vn; vn = tmp;
        ++n;
        if (max_sweeps > 0 && n >= max_sweeps) break;
    }
    memcpy(v_out, v, sizeof(double) * S);
    if (n_out) *n_out = n;
    free(v); free(vn);
    return 0;
}

/* B independent worlds, one per OpenMP thread: compute_expected_svf (maxent.py:162-193) of each.
 * Tables of world b start at b * S*A*K. */
int oracle_batch_maxent_step(int B, int S, int A, int K, const int *sidx, const double *sp,
                             const unsigned char *term, const double *p0, const double *reward, double eps,
                             long max_sweeps, double *svf, long *n_svf) {
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const size_t to = (size_t)b * S * A * K;
        double *pol = malloc(sizeof(double) * (size_t)S * A);
        if (!pol) { rc = -1; continue; }
        if (oracle_backward(S, A, K, sidx + to, sp + to, term, reward + (size_t)b * S, 2L * S, 0, pol)) rc = -1;
        if (oracle_svf(S, A, K, sidx + to, sp + to, p0, term, pol, eps, max_sweeps, svf + (size_t)b * S, n_svf + b)) rc = -1;
        free(pol);
    }
    return rc;
}
