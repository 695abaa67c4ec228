// dense_batch.cu -- the batched DENSE path: B reward candidates that share one dense p_transition.
//
// The reference's arithmetic is a dense product per action and sweep (`p[a].dot(x)`, maxent.py:155 /
// :329; `p[a].T.dot(p_action[:, a] * d)`, :109).  For tables that really are dense (K ~ S successors per
// state) and a batch of B candidates sharing the table, a sweep over the whole batch IS a matrix
// product [rows x S] . [S x B] (BASELINE north_star (4), configs[3]), and the ELL gather is the wrong
// tool: it re-reads index + value per entry for every candidate and shares nothing between them.
//
//   irlb200_dense_pack     one pass over P[S][S'][A] (A innermost, gridworld.py:124-142) -> three row-major
//                          matrices with the contraction index contiguous:
//                            Pa [A*S][S]   Pa[a*S + s][s'] = P[s, s', a]        soft-VI / VI / last backward sweep
//                            Pm [S][S]     Pm[s][s'] = sum_a P[s, s', a]         backward sweeps (merged weights)
//                            Pt [S][A*S]   Pt[s'][a*S + s] = P[s, s', a]         forward pass (transposed)
//   dense_gemm_kernel      C^T[n][m] = sum_k A[m][k] * X^T[n][k]: FP64 tensor-core contraction
//                          (mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 path), 64 x 64 CTA tiles, 4 warps of
//                          4 x 4 MMA tiles, operands staged in shared memory by a 3-stage cp.async pipeline
//                          in k-blocks of four ([kb][row][4]: a fragment load of a warp is 256 contiguous
//                          bytes, no bank conflicts); iterates are candidate-major ([B][S], the layout of the
//                          rewards), which is exactly the "col" operand of the MMA.
//   epilogues              per (candidate, state) the reference's elementwise arithmetic on the contraction's
//                          output -- softmax fold / max (succ_update of phases.cuh, so the same expression
//                          tree as every other kernel), er * dot, p_initial + sum -- plus the stop rule per
//                          candidate (`while delta > eps`), each candidate frozen at ITS OWN stopping sweep.
//
// One GEMM launch + one epilogue launch + one B-thread bookkeeping launch per sweep: at S = 1024, B = 4096 a
// sweep is 8.6 - 34 GFLOP, so launch latency is noise and the host only polls the number of live candidates
// every few sweeps.  Summation order inside a dot product differs from a BLAS dgemv (as the ELL kernels'
// does); the parity tests hold the results to 1e-10 of the reference arithmetic with identical sweep counts.
#include <cstdlib>

#include "host_util.h"
#include "phases.cuh"

namespace irlb200 {

// ---------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------
__global__ void dense_pack_kernel(const double *__restrict__ P, int S, int A, double *__restrict__ Pa,
                                  double *__restrict__ Pm, double *__restrict__ Pt) {
    // one thread per (s, s'): reads the A contiguous values P[s][s'][:]
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)S * S) return;
    const int s = (int)(i / S), sp = (int)(i % S);
    double sum = 0.0;
    for (int a = 0; a < A; ++a) {
        const double v = P[i * A + a];
        Pa[((size_t)a * S + s) * S + sp] = v;
        Pt[(size_t)sp * A * S + (size_t)a * S + s] = v;
        sum = a == 0 ? v : sum + v;
    }
    Pm[i] = sum;
}

// ---------------------------------------------------------------------------
// FP64 tensor-core GEMM:  Ct[n][m] = sum_k A[m][k] * Xt[n][k]     (A: M x K row-major, Xt: N x K row-major)
// ---------------------------------------------------------------------------
constexpr int kTK = 16, kStages = 3;

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool pred) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = pred ? 8 : 0;                                // src-size 0: zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// the sm_90+ FP64 MMA shape: 16 x 8 x 16, 2 048 FMA per warp instruction (8 x the m8n8k4 of sm_80).  Fragments (CuTe
// MMA_Traits<SM90_16x8x16_F64F64F64F64_TN>): a_i: row = lane / 4 + 8 (i % 2), k = lane % 4 + 4 (i / 2);
// b_v: k = lane % 4 + 4 v, n = lane / 4;  c_i: row = lane / 4 + 8 (i / 2), col = 2 (lane % 4) + i % 2.
__device__ __forceinline__ void dmma_m16n8k16(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7, %8, %9, %10, %11}, "
                 "{%12, %13, %14, %15}, {%0, %1, %2, %3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// CTA tile = (32 WM) x (32 WN), one warp per 32 x 32 sub-tile (4 x 4 MMA tiles, 32 accumulators per thread).
//   <2, 2>: 64 x 64, 128 threads, 4 CTAs per SM   -- small problems (many CTAs)
//   <4, 4>: 128 x 128, 512 threads, 1 CTA per SM  -- 16 flop per byte fetched from L2 instead of 8: the 64 x 64
//           tile moves ~69 B per clock and SM from L2 at full MMA rate, which is what bounds it
template <int WM, int WN>
struct GemmCfg {
    static constexpr int TM = 32 * WM, TN = 32 * WN, THREADS = 32 * WM * WN;
    // shared layout of one operand tile: [k-block of 4][row][4], k-block stride padded by 4 doubles so that the
    // four k-blocks a copied row lands in start 8 banks apart
    static constexpr int KBS_A = TM * 4 + 4, KBS_X = TN * 4 + 4;
    static constexpr int A_DOUBLES = (kTK / 4) * KBS_A, X_DOUBLES = (kTK / 4) * KBS_X;
    static constexpr int STAGE_DOUBLES = A_DOUBLES + X_DOUBLES;
    static constexpr size_t SMEM = sizeof(double) * kStages * STAGE_DOUBLES;
};

// rows [row0, row0 + ROWS) x k [k0, k0 + 16) of a row-major matrix (ld = K) -> smem [kb 0..3][row][4]
template <int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(double *dst, const double *__restrict__ src, int rows, int K, int row0,
                                          int k0, int tid) {
    constexpr int KBS = ROWS * 4 + 4;
    if ((K & 1) == 0) {
        // even K: rows start 16-byte aligned, two consecutive k travel as one 16-byte cp.async
#pragma unroll
        for (int it = 0; it < (ROWS * kTK / 2) / THREADS; ++it) {
            const int e = it * THREADS + tid;                      // pair index of the tile, k fastest
            const int r = e / (kTK / 2), k = (e % (kTK / 2)) * 2;
            const bool ok = (row0 + r) < rows && (k0 + k) < K;     // K even: the pair is in or out as a whole
            const double *g = src + (ok ? ((size_t)(row0 + r) * K + k0 + k) : 0);
            cp_async16(dst + (k >> 2) * KBS + r * 4 + (k & 3), g, ok ? 16 : 0);
        }
        return;
    }
#pragma unroll
    for (int it = 0; it < (ROWS * kTK) / THREADS; ++it) {
        const int e = it * THREADS + tid;                          // element of the tile, k fastest
        const int r = e / kTK, k = e % kTK;
        const bool ok = (row0 + r) < rows && (k0 + k) < K;
        const double *g = src + (ok ? ((size_t)(row0 + r) * K + k0 + k) : 0);
        cp_async8(dst + (k >> 2) * KBS + r * 4 + (k & 3), g, ok);
    }
}

template <int WM, int WN>
__global__ void __launch_bounds__(32 * WM * WN, (WM * WN <= 4) ? 4 : 1)
    dense_gemm_kernel(const double *__restrict__ Am, const double *__restrict__ Xt, double *__restrict__ Ct, int M,
                      int N, int K) {
    using Cfg = GemmCfg<WM, WN>;
    extern __shared__ __align__(16) double smem_d[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int m0 = blockIdx.x * Cfg::TM, n0 = blockIdx.y * Cfg::TN;
    const int nk = (K + kTK - 1) / kTK;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto stage = [&](int s) { return smem_d + (size_t)s * Cfg::STAGE_DOUBLES; };
    auto issue = [&](int kc) {
        if (kc < nk) {
            double *st = stage(kc % kStages);
            load_tile<Cfg::TM, Cfg::THREADS>(st, Am, M, K, m0, kc * kTK, tid);
            load_tile<Cfg::TN, Cfg::THREADS>(st + Cfg::A_DOUBLES, Xt, N, K, n0, kc * kTK, tid);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    for (int kc = 0; kc < nk; ++kc) {
        issue(kc + 2);
        cp_async_wait<2>();
        __syncthreads();
        const double *As = stage(kc % kStages), *Xs = As + Cfg::A_DOUBLES;
#pragma unroll
        for (int kb = 0; kb < kTK / 4; ++kb) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[kb * Cfg::KBS_A + (wm * 32 + i * 8) * 4 + lane];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Xs[kb * Cfg::KBS_X + (wn * 32 + j * 8) * 4 + lane];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    // C fragment: row m = lane >> 2, columns n = 2 * (lane & 3) + {0, 1}; stored candidate-major Ct[n][m]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm * 32 + i * 8 + (lane >> 2);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
            if (n < N) Ct[(size_t)n * M + m] = acc[i][j][0];
            if (n + 1 < N) Ct[(size_t)(n + 1) * M + m] = acc[i][j][1];
        }
    }
}

// The same tiling with the 16 x 8 x 16 MMA: per k-chunk of 16 a warp issues 2 x 4 MMAs instead of 64.
template <int WM, int WN, int MINB>
__global__ void __launch_bounds__(32 * WM * WN, MINB)
    dense_gemm16_kernel(const double *__restrict__ Am, const double *__restrict__ Xt, double *__restrict__ Ct, int M,
                        int N, int K) {
    using Cfg = GemmCfg<WM, WN>;
    extern __shared__ __align__(16) double smem_d[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int m0 = blockIdx.x * Cfg::TM, n0 = blockIdx.y * Cfg::TN;
    const int nk = (K + kTK - 1) / kTK;
    double acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.0;
    auto stage = [&](int s) { return smem_d + (size_t)s * Cfg::STAGE_DOUBLES; };
    auto issue = [&](int kc) {
        if (kc < nk) {
            double *st = stage(kc % kStages);
            load_tile<Cfg::TM, Cfg::THREADS>(st, Am, M, K, m0, kc * kTK, tid);
            load_tile<Cfg::TN, Cfg::THREADS>(st + Cfg::A_DOUBLES, Xt, N, K, n0, kc * kTK, tid);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    for (int kc = 0; kc < nk; ++kc) {
        issue(kc + 2);
        cp_async_wait<2>();
        __syncthreads();
        const double *As = stage(kc % kStages), *Xs = As + Cfg::A_DOUBLES;
        double bf[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) bf[j][v] = Xs[v * Cfg::KBS_X + (wn * 32 + j * 8) * 4 + lane];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double af[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) af[v] = As[(v >> 1) * Cfg::KBS_A + (wm * 32 + i * 16 + 8 * (v & 1)) * 4 + lane];
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_m16n8k16(acc[i][j], af, bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int m = m0 + wm * 32 + i * 16 + (lane >> 2) + 8 * (v >> 1);
                const int n = n0 + wn * 32 + j * 8 + 2 * (lane & 3) + (v & 1);
                if (m < M && n < N) Ct[(size_t)n * M + m] = acc[i][j][v];
            }
}

template <int WM, int WN, int MINB>
static int launch_gemm16_cfg(const double *Am, const double *Xt, double *Ct, int M, int N, int K, cudaStream_t st) {
    using Cfg = GemmCfg<WM, WN>;
    if (Cfg::SMEM > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dense_gemm16_kernel<WM, WN, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(dense gemm16)");
    }
    dim3 grid((M + Cfg::TM - 1) / Cfg::TM, (N + Cfg::TN - 1) / Cfg::TN);
    dense_gemm16_kernel<WM, WN, MINB><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(Am, Xt, Ct, M, N, K);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "dense_gemm16_kernel");
}

template <int WM, int WN>
static int launch_gemm_cfg(const double *Am, const double *Xt, double *Ct, int M, int N, int K, cudaStream_t st) {
    using Cfg = GemmCfg<WM, WN>;
    if (Cfg::SMEM > 48 * 1024) {          // per device and context: set on every call (cheap), like the other launchers
        cudaError_t e = cudaFuncSetAttribute(dense_gemm_kernel<WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(dense gemm)");
    }
    dim3 grid((M + Cfg::TM - 1) / Cfg::TM, (N + Cfg::TN - 1) / Cfg::TN);
    dense_gemm_kernel<WM, WN><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(Am, Xt, Ct, M, N, K);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "dense_gemm_kernel");
}

static int launch_gemm(const double *Am, const double *Xt, double *Ct, int M, int N, int K, cudaStream_t st) {
    static int force = -1;
    if (force < 0) {
        const char *v = getenv("IRLB200_DENSE_TILE");              // 128: the 128 x 128 tile (experiments)
        force = v ? atoi(v) : 0;
    }
    // measured at S = 1024, A = 4, B = 4096 (soft-VI sweeps): 64 x 64 tiles 18.4 TFLOP/s, 128 x 128 tiles 17.1 --
    // the small tile's four CTAs per SM hide more latency than the big tile saves in L2 traffic
    if (force == 128) return launch_gemm_cfg<4, 4>(Am, Xt, Ct, M, N, K, st);
    if (force == 64) return launch_gemm_cfg<2, 2>(Am, Xt, Ct, M, N, K, st);          // m8n8k4 (sm_80 shape)
    if (force == 1616) return launch_gemm16_cfg<4, 4, 1>(Am, Xt, Ct, M, N, K, st);   // m16n8k16, 128 x 128 tile
    if (force == 164) return launch_gemm16_cfg<2, 2, 4>(Am, Xt, Ct, M, N, K, st);    // 64 x 64, 4 CTAs per SM (128 registers)
    if (force == 1642) return launch_gemm16_cfg<4, 2, 2>(Am, Xt, Ct, M, N, K, st);   // 128 x 64, 8 warps, 2 CTAs per SM
    if (force == 1624) return launch_gemm16_cfg<2, 4, 2>(Am, Xt, Ct, M, N, K, st);   // 64 x 128
    return launch_gemm16_cfg<2, 2, 3>(Am, Xt, Ct, M, N, K, st);                      // m16n8k16, 64 x 64 tile
}

// ---------------------------------------------------------------------------
// per-candidate loop state
// ---------------------------------------------------------------------------
struct DenseCtl {
    unsigned long long *delta_bits;   // [B] max |diff| of the current sweep as an ordered bit pattern
    int *nan;                         // [B]
    int *active;                      // [B]
    int32_t *n_iter, *status;         // [B]
    double *colmax;                   // [B] backward: column maximum of the current sweep
    double *scale;                    // [B] backward: power of two applied to the next sweep
    int *n_active;                    // [1]
    int *done_now;                    // [B] 1: stopped in the sweep just finalised (consumed by the policy kernel)
};

// max over the warp of a non-negative double (its bit pattern orders like the value)
__device__ __forceinline__ double warp_max_nonneg(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// One atomic per warp instead of one per state when the warp's lanes belong to one candidate (always, when
// S % 32 == 0): S atomics on ONE address per candidate and sweep serialise in L2.  Must be called by full warps.
__device__ __forceinline__ void vote_diff(const DenseCtl &c, int b, double x_new, double x_old, bool live) {
    const double diff = live ? fabs(x_new - x_old) : 0.0;
    const bool nan = diff != diff;
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    if (__all_sync(0xffffffffu, b == b0 || !live)) {
        const bool any_nan = __any_sync(0xffffffffu, nan);
        const double m = warp_max_nonneg(nan ? 0.0 : diff);
        if ((threadIdx.x & 31) == 0) {
            if (any_nan) c.nan[b0] = 1;
            atomicMax(c.delta_bits + b0, (unsigned long long)__double_as_longlong(m));
        }
    } else if (live) {
        if (nan) c.nan[b] = 1;
        else atomicMax(c.delta_bits + b, (unsigned long long)__double_as_longlong(diff));
    }
}

// after every sweep, one thread per candidate: `while delta > eps` (NaN ends the loop), max-sweep guard
__global__ void dense_finalize_kernel(DenseCtl c, int B, double eps, int max_sweeps) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !c.active[b]) return;
    const int n = ++c.n_iter[b];
    const double d = __longlong_as_double((long long)c.delta_bits[b]);
    int st = -1;
    if (c.nan[b]) st = IRLB200_ST_NONFINITE;
    else if (!(d > eps)) st = IRLB200_ST_CONVERGED;
    else if (max_sweeps > 0 && n >= max_sweeps) st = IRLB200_ST_MAXSWEEPS;
    c.delta_bits[b] = 0ull;
    c.nan[b] = 0;
    if (st >= 0) {
        c.status[b] = st;
        c.active[b] = 0;
        c.done_now[b] = 1;
        atomicSub(c.n_active, 1);
    }
}

__global__ void dense_init_ctl_kernel(DenseCtl c, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *c.n_active = B;
    if (b >= B) return;
    c.delta_bits[b] = 0ull; c.nan[b] = 0; c.active[b] = 1; c.n_iter[b] = 0; c.status[b] = IRLB200_ST_CONVERGED;
    c.colmax[b] = 0.0; c.scale[b] = 1.0; c.done_now[b] = 0;
}

// ---------------------------------------------------------------------------
// epilogues (one thread per (candidate, state); T = output of the contraction, candidate-major)
// ---------------------------------------------------------------------------
// soft value iteration (maxent.py:329-341) / value iteration (solver.py:44-50): T[b][a*S + s] = P_a[s,:] . x_b
// q_a and the new value of one (candidate, state) from the contraction's output: the expression tree of
// phases.cuh::succ_update with dot(a) = t[a * S]
template <int OP>
__device__ __forceinline__ double dense_succ_value(const double *t, int S, int A, double r, double c1, double discount,
                                                   int vi_mean, double *q) {
    for (int a = 0; a < A; ++a) {
        const double dot = t[(size_t)a * S];
        q[a] = (OP == kOpSoftVI) ? r + discount * dot : discount * dot;
    }
    double xn;
    if (OP == kOpSoftVI) {
        double m = c1;
        for (int a = 0; a < A; ++a) m = max_nan(m, q[a]);
        if (fabs(m) < INFINITY) {
            double ssum = 0.0;
            if (c1 != -INFINITY) ssum = exp(c1 - m);
            for (int a = 0; a < A; ++a) ssum += exp(q[a] - m);
            xn = m + log(ssum);
        } else {
            xn = c1;
            for (int a = 0; a < A; ++a) xn = softmax2(xn, q[a]);
        }
    } else if (vi_mean) {
        xn = q[0];
        for (int a = 1; a < A; ++a) xn += q[a];
        xn = r + xn / (double)A;
    } else {
        xn = q[0];
        for (int a = 1; a < A; ++a) xn = max_nan(xn, q[a]);
        xn = r + xn;
    }
    return xn;
}

template <int OP>
__global__ void __launch_bounds__(256)
    dense_succ_epilogue(const double *__restrict__ T, const double *__restrict__ reward, const double *__restrict__ phi,
                        double *__restrict__ X, DenseCtl c, int S, int A, int B, double discount, int vi_mean) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < (size_t)B * S;
    const int b = in ? (int)(i / S) : 0, s = in ? (int)(i % S) : 0;
    const bool live = in && c.active[b];
    double xn = 0.0, xo = 0.0;
    if (live) {
        double q[kMaxDynA];
        xn = dense_succ_value<OP>(T + (size_t)b * A * S + s, S, A, reward[i], (OP == kOpSoftVI) ? phi[s] : 0.0, discount,
                                  vi_mean, q);
        xo = X[i];
        X[i] = xn;
    }
    vote_diff(c, b, xn, xo, live);
}

// policy of the candidates that stopped in the sweep just finalised (maxent.py:341): q of that sweep (T is still
// the contraction's output of that sweep) against the value it produced
__global__ void __launch_bounds__(256)
    dense_policy_kernel(const double *__restrict__ T, const double *__restrict__ reward, const double *__restrict__ X,
                        double *__restrict__ policy, DenseCtl c, int S, int A, int B, double discount) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * S) return;
    const int b = (int)(i / S), s = (int)(i % S);
    if (!c.done_now[b]) return;
    const double r = reward[i], x = X[i];
    const double *t = T + (size_t)b * A * S + s;
    for (int a = 0; a < A; ++a) policy[i * A + a] = exp(r + discount * t[(size_t)a * S] - x);
}
__global__ void dense_clear_done_kernel(DenseCtl c, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) c.done_now[b] = 0;
}

// backward sweeps 1 .. n-1 (merged weights): zs' = er * (Pm . zs), times the candidate's power-of-two scale
__global__ void dense_backward_epilogue(const double *__restrict__ T, const double *__restrict__ reward,
                                        double *__restrict__ X, DenseCtl c, int S, int B) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < (size_t)B * S;
    const int b = in ? (int)(i / S) : 0;
    double z = 0.0;
    if (in) {
        z = exp(reward[i]) * T[i] * c.scale[b];
        X[i] = z;
    }
    const double zz = (in && z == z) ? fmax(z, 0.0) : 0.0;
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    if (__all_sync(0xffffffffu, b == b0 || !in)) {
        const double m = warp_max_nonneg(zz);
        if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)c.colmax + b0, (unsigned long long)__double_as_longlong(m));
    } else if (in) {
        atomicMax((unsigned long long *)c.colmax + b, (unsigned long long)__double_as_longlong(zz));
    }
}
__global__ void dense_backward_rescale(DenseCtl c, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double m = c.colmax[b];
    c.scale[b] = (m > 0.0 && m < INFINITY) ? scale_pow2(1.0, -frexp_exponent(m)) : 1.0;
    c.colmax[b] = 0.0;
}
// last backward sweep, per action exactly as maxent.py:155-159: za = er * P_a.dot(zs), policy = za / sum_a za
__global__ void dense_backward_last(const double *__restrict__ T, const double *__restrict__ reward,
                                    double *__restrict__ policy, int S, int A, int B) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * S) return;
    const int b = (int)(i / S), s = (int)(i % S);
    const double er = exp(reward[i]);
    const double *t = T + (size_t)b * A * S + s;
    double za[kMaxDynA], zs = 0.0;
    for (int a = 0; a < A; ++a) {
        za[a] = er * t[(size_t)a * S];
        zs = a == 0 ? za[0] : zs + za[a];
    }
    for (int a = 0; a < A; ++a) policy[i * A + a] = za[a] / zs;
}
__global__ void dense_backward_init(const uint8_t *__restrict__ term, double *__restrict__ X, int S, int B) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * S) return;
    X[i] = term[i % S] ? 1.0 : 0.0;                                                        // maxent.py:146-147
}

// forward pass (maxent.py:98-112): Y[b][a*S + s] = policy[b,s,a] * d[b,s], 0 for terminal s (their rows are dropped)
__global__ void dense_svf_prep(const double *__restrict__ policy, const double *__restrict__ X,
                               const uint8_t *__restrict__ term, double *__restrict__ Y, DenseCtl c, int S, int A, int B) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * S) return;
    const int b = (int)(i / S), s = (int)(i % S);
    if (!c.active[b]) return;
    const double d = X[i];
    const bool t = term[s] != 0;
    for (int a = 0; a < A; ++a) Y[(size_t)b * A * S + (size_t)a * S + s] = t ? 0.0 : policy[i * A + a] * d;
}
__global__ void dense_svf_epilogue(const double *__restrict__ T, const double *__restrict__ p0, size_t p0_stride,
                                   double *__restrict__ X, DenseCtl c, int S, int B) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < (size_t)B * S;
    const int b = in ? (int)(i / S) : 0, s = in ? (int)(i % S) : 0;
    const bool live = in && c.active[b];
    double xn = 0.0, xo = 0.0;
    if (live) {
        xn = p0[(size_t)b * p0_stride + s] + T[i];                                          // :110
        xo = X[i];
        X[i] = xn;
    }
    vote_diff(c, b, xn, xo, live);
}
__global__ void dense_fill_kernel(double *X, size_t n, double v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) X[i] = v;
}
__global__ void dense_grad_kernel(const double *__restrict__ svf, const double *__restrict__ ef, size_t ef_stride,
                                  double *__restrict__ grad, int S, int B) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * S) return;
    grad[i] = ef[(i / S) * ef_stride + i % S] - svf[i];                                     // :248, identity features
}

struct DenseWork {
    double *T, *Y;
    DenseCtl c;
};
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t dense_work_bytes_impl(int S, int A, int B) {
    return 2 * align256(sizeof(double) * (size_t)B * A * S) + 11 * align256(sizeof(double) * (size_t)B) + 256;
}
static void carve_work(DenseWork &w, void *work, int S, int A, int B) {
    unsigned char *p = static_cast<unsigned char *>(work);
    auto take = [&](size_t bytes) { void *r = p; p += align256(bytes); return r; };
    w.T = static_cast<double *>(take(sizeof(double) * (size_t)B * A * S));
    w.Y = static_cast<double *>(take(sizeof(double) * (size_t)B * A * S));
    w.c.delta_bits = static_cast<unsigned long long *>(take(8 * (size_t)B));
    w.c.nan = static_cast<int *>(take(8 * (size_t)B));
    w.c.active = static_cast<int *>(take(8 * (size_t)B));
    w.c.colmax = static_cast<double *>(take(8 * (size_t)B));
    w.c.scale = static_cast<double *>(take(8 * (size_t)B));
    w.c.done_now = static_cast<int *>(take(8 * (size_t)B));
    w.c.n_iter = static_cast<int32_t *>(take(8 * (size_t)B));          // callers with their own outputs override these
    w.c.status = static_cast<int32_t *>(take(8 * (size_t)B));
    w.c.n_active = static_cast<int *>(take(256));
}

// run sweeps until no candidate is live; the host looks at the live count every `poll` sweeps
template <class Sweep, class After>
static int run_until_done(Sweep sweep, After after, DenseCtl c, int B, double eps, int max_sweeps, cudaStream_t st) {
    const int poll = 8;
    int live = B;
    for (long long done = 0; live > 0; done += poll) {
        for (int i = 0; i < poll; ++i) {
            if (int rc = sweep()) return rc;
            dense_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(c, B, eps, max_sweeps);
            after();                                         // what only the candidates that just stopped need
        }
        cudaError_t e = cudaMemcpyAsync(&live, c.n_active, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail_cuda(e, "dense batch: sweep loop");
    }
    return IRLB200_OK;
}

}  // namespace irlb200

using namespace irlb200;

extern "C" size_t irlb200_dense_pack_doubles(int S, int A) { return (size_t)S * S * (2 * (size_t)A + 1); }
extern "C" size_t irlb200_dense_batch_work_bytes(int S, int A, int B) { return dense_work_bytes_impl(S, A, B); }

extern "C" int irlb200_dense_pack(const double *P, int S, int A, double *packed, void *stream) {
    if (!P || !packed || S <= 0 || A <= 0) return fail(IRLB200_EINVAL, "dense_pack: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    double *Pa = packed, *Pm = Pa + (size_t)A * S * S, *Pt = Pm + (size_t)S * S;
    const size_t n = (size_t)S * S;
    dense_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, S, A, Pa, Pm, Pt);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "dense_pack_kernel");
}

static int dense_common_check(const double *packed, int S, int A, int B, void *work, size_t work_bytes) {
    if (!packed || S <= 0 || A <= 0 || B <= 0 || !work) return fail(IRLB200_EINVAL, "dense batch: bad argument");
    if (A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
    if (work_bytes < dense_work_bytes_impl(S, A, B)) return fail(IRLB200_EINVAL, "dense batch: work buffer too small");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    return IRLB200_OK;
}

// local_action_probabilities (maxent.py:119-159) for B candidates sharing the dense table
extern "C" int irlb200_dense_batch_backward(const double *packed, int S, int A, int B, const double *reward,
                                            const uint8_t *terminal_mask, int n_sweeps, double *policy, double *iterate,
                                            void *work, size_t work_bytes, void *stream) {
    if (int rc = dense_common_check(packed, S, A, B, work, work_bytes)) return rc;
    if (!reward || !terminal_mask || !policy || !iterate || n_sweeps < 1) return fail(IRLB200_EINVAL, "dense_batch_backward: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const double *Pa = packed, *Pm = Pa + (size_t)A * S * S;
    DenseWork w;
    carve_work(w, work, S, A, B);
    const size_t n = (size_t)B * S;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    dense_init_ctl_kernel<<<(B + 127) / 128, 128, 0, st>>>(w.c, B);
    dense_backward_init<<<blocks, 256, 0, st>>>(terminal_mask, iterate, S, B);
    for (int t = 0; t + 1 < n_sweeps; ++t) {
        if (int rc = launch_gemm(Pm, iterate, w.T, S, B, S, st)) return rc;
        dense_backward_epilogue<<<blocks, 256, 0, st>>>(w.T, reward, iterate, w.c, S, B);
        dense_backward_rescale<<<(B + 127) / 128, 128, 0, st>>>(w.c, B);
    }
    if (int rc = launch_gemm(Pa, iterate, w.T, A * S, B, S, st)) return rc;
    dense_backward_last<<<blocks, 256, 0, st>>>(w.T, reward, policy, S, A, B);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "dense_batch_backward");
}

// op 1: local_causal_action_probabilities (maxent.py:279-341); op 2: value_iteration (solver.py:9-52, kind 1: mean)
extern "C" int irlb200_dense_batch_succ(int op, const double *packed, int S, int A, int B, const double *reward,
                                        const double *phi, double discount, double eps, int max_sweeps, int vi_mean,
                                        double *value, double *policy, int32_t *n_iter, int32_t *status, void *work,
                                        size_t work_bytes, void *stream) {
    if (int rc = dense_common_check(packed, S, A, B, work, work_bytes)) return rc;
    if ((op != 1 && op != 2) || !reward || !value || !n_iter || !status || (op == 1 && (!phi || !policy)))
        return fail(IRLB200_EINVAL, "dense_batch_succ: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const double *Pa = packed;
    DenseWork w;
    carve_work(w, work, S, A, B);
    w.c.n_iter = n_iter; w.c.status = status;
    const size_t n = (size_t)B * S;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    dense_init_ctl_kernel<<<(B + 127) / 128, 128, 0, st>>>(w.c, B);
    dense_fill_kernel<<<blocks, 256, 0, st>>>(value, n, op == 1 ? kNegHuge : 0.0);           // maxent.py:323 / solver.py:29
    auto sweep = [&]() -> int {
        if (int rc = launch_gemm(Pa, value, w.T, A * S, B, S, st)) return rc;
        if (op == 1) dense_succ_epilogue<kOpSoftVI><<<blocks, 256, 0, st>>>(w.T, reward, phi, value, w.c, S, A, B, discount, 0);
        else dense_succ_epilogue<kOpVI><<<blocks, 256, 0, st>>>(w.T, reward, nullptr, value, w.c, S, A, B, discount, vi_mean);
        return IRLB200_OK;
    };
    auto after = [&]() {
        if (op != 1) return;
        dense_policy_kernel<<<blocks, 256, 0, st>>>(w.T, reward, value, policy, w.c, S, A, B, discount);
        dense_clear_done_kernel<<<(B + 127) / 128, 128, 0, st>>>(w.c, B);
    };
    return run_until_done(sweep, after, w.c, B, eps, max_sweeps, st);
}

// expected_svf_from_policy (maxent.py:63-114) for B policies over the shared dense table
extern "C" int irlb200_dense_batch_svf(const double *packed, int S, int A, int B, const double *p_initial, int p0_shared,
                                       const uint8_t *terminal_mask, const double *policy, double eps, int max_sweeps,
                                       double *svf, const double *e_features, int ef_shared, double *grad,
                                       int32_t *n_iter, int32_t *status, void *work, size_t work_bytes, void *stream) {
    if (int rc = dense_common_check(packed, S, A, B, work, work_bytes)) return rc;
    if (!p_initial || !terminal_mask || !policy || !svf || !n_iter || !status || (grad && !e_features))
        return fail(IRLB200_EINVAL, "dense_batch_svf: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const double *Pt = packed + (size_t)A * S * S + (size_t)S * S;
    DenseWork w;
    carve_work(w, work, S, A, B);
    w.c.n_iter = n_iter; w.c.status = status;
    const size_t n = (size_t)B * S;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    dense_init_ctl_kernel<<<(B + 127) / 128, 128, 0, st>>>(w.c, B);
    dense_fill_kernel<<<blocks, 256, 0, st>>>(svf, n, 0.0);                                   // d = np.zeros  :105
    auto sweep = [&]() -> int {
        dense_svf_prep<<<blocks, 256, 0, st>>>(policy, svf, terminal_mask, w.Y, w.c, S, A, B);
        if (int rc = launch_gemm(Pt, w.Y, w.T, S, B, A * S, st)) return rc;
        dense_svf_epilogue<<<blocks, 256, 0, st>>>(w.T, p_initial, p0_shared ? 0 : (size_t)S, svf, w.c, S, B);
        return IRLB200_OK;
    };
    if (int rc = run_until_done(sweep, [] {}, w.c, B, eps, max_sweeps, st)) return rc;
    if (grad) dense_grad_kernel<<<blocks, 256, 0, st>>>(svf, e_features, ef_shared ? 0 : (size_t)S, grad, S, B);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "dense_batch_svf");
}
