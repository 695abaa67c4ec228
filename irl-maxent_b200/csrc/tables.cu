// tables.cu -- kernel (1): building the state-merged ELL tables.
//
//   dense_count / dense_fill   one-time compression of the reference's dense
//                              p_transition[S][S'][A] (replaces the per-call
//                              slicing at maxent.py:102,143,320, solver.py:37)
//   gridworld_tables           the same tables for GridWorld / IcyGridWorld
//                              straight from (size, p_slip), no dense detour
//   gridworld_dense            dense table of those worlds, on the device
//   features_dot / _grad       the two dense feature products of the outer loop
#include "common.cuh"
#include "host_util.h"

namespace irlb200 {

// ---------------------------------------------------------------------------
// dense -> tables
// ---------------------------------------------------------------------------
// One CTA per source state s: the row P[s][:][:] is S*A contiguous doubles.
template <bool FILL>
__global__ void __launch_bounds__(256)
dense_rows_kernel(const double *__restrict__ P, int S, int A, int Ks, int Kp,
                  int32_t *succ_cnt, int32_t *pred_cnt, int32_t *kmax,
                  int32_t *succ_idx, double *succ_p, int32_t *pred_idx, double *pred_p,
                  int32_t *pred_cursor) {
    const int s = blockIdx.x;
    const double *row = P + (size_t)s * S * A;
    __shared__ int warp_tot[8];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int t0 = 0; t0 < S; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        bool nz = false;
        if (t < S) {
            for (int a = 0; a < A; ++a) nz |= (row[(size_t)t * A + a] != 0.0);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = base;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        const int j = before + __popc(bal & ((1u << lane) - 1u));
        if (nz) {
            if (FILL) {
                succ_idx[(size_t)j * S + s] = t;
                for (int a = 0; a < A; ++a)
                    succ_p[((size_t)a * Ks + j) * S + s] = row[(size_t)t * A + a];
                const int jp = atomicAdd(&pred_cursor[t], 1);
                pred_idx[(size_t)jp * S + t] = s;
                for (int a = 0; a < A; ++a)
                    pred_p[((size_t)a * Kp + jp) * S + t] = row[(size_t)t * A + a];
            } else {
                atomicAdd(&pred_cnt[t], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_tot[w];
            base += tot;
        }
        __syncthreads();
    }
    const int count = base;
    if (FILL) {
        for (int j = count + threadIdx.x; j < Ks; j += blockDim.x) {      // padding slots
            succ_idx[(size_t)j * S + s] = s;
            for (int a = 0; a < A; ++a) succ_p[((size_t)a * Ks + j) * S + s] = 0.0;
        }
    } else if (threadIdx.x == 0) {
        succ_cnt[s] = count;
        atomicMax(&kmax[0], count);
    }
}

__global__ void max_count_kernel(const int32_t *cnt, int S, int32_t *out) {
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) m = max(m, cnt[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// The predecessor slots were claimed with atomics (arbitrary order): sort every
// state's slots by predecessor index and pad the tail.  One thread per state.
__global__ void pred_sort_kernel(int S, int A, int Kp, const int32_t *pred_cnt,
                                 int32_t *pred_idx, double *pred_p) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int n = pred_cnt[s];
    for (int i = 1; i < n; ++i) {
        for (int j = i; j > 0; --j) {
            int32_t *x = pred_idx + (size_t)j * S + s, *y = pred_idx + (size_t)(j - 1) * S + s;
            if (*y <= *x) break;
            int32_t ti = *x; *x = *y; *y = ti;
            for (int a = 0; a < A; ++a) {
                double *u = pred_p + ((size_t)a * Kp + j) * S + s, *v = pred_p + ((size_t)a * Kp + j - 1) * S + s;
                double td = *u; *u = *v; *v = td;
            }
        }
    }
    for (int j = n; j < Kp; ++j) {
        pred_idx[(size_t)j * S + s] = s;
        for (int a = 0; a < A; ++a) pred_p[((size_t)a * Kp + j) * S + s] = 0.0;
    }
}

// ---------------------------------------------------------------------------
// grid worlds
// ---------------------------------------------------------------------------
// actions as coordinate offsets, reference order (gridworld.py:47)
__device__ __constant__ int kActX[4] = {1, -1, 0, 0};
__device__ __constant__ int kActY[4] = {0, 0, 1, -1};

// P[s_from, s_to, a] of GridWorld (icy == 0, gridworld.py:144-171) or IcyGridWorld
// (gridworld.py:200-248).  Every value is formed with the same IEEE operations,
// in the same order, as the reference's Python expressions (no FMA contraction:
// explicit _rn intrinsics), so tables are bit-identical.
__device__ double world_prob(int n, int icy, double p, int s_from, int s_to, int a) {
    const int fx = s_from % n, fy = s_from / n, tx = s_to % n, ty = s_to / n;
    const int gx = fx + kActX[a], gy = fy + kActY[a];
    const bool inside = gx >= 0 && gx < n && gy >= 0 && gy < n;
    const bool same = fx == tx && fy == ty;
    if (!icy) {
        if (gx == tx && gy == ty) return 1.0;
        return (same && !inside) ? 1.0 : 0.0;
    }
    const double nA = 4.0;
    const double keep = __dadd_rn(1.0, -p);                                   // 1.0 - p_slip
    if (gx == tx && gy == ty) return __dadd_rn(keep, __ddiv_rn(p, nA));       // :219
    if (abs(fx - tx) + abs(fy - ty) == 1) return __ddiv_rn(p, nA);            // :223
    if (!same) return 0.0;
    const bool xb = !(fx > 0 && fx < n - 1), yb = !(fy > 0 && fy < n - 1);
    const double two_p = __dmul_rn(2.0, p);
    if (!inside) {
        if (xb && yb) return __dadd_rn(keep, __ddiv_rn(two_p, nA));           // :231
        return __dadd_rn(keep, __ddiv_rn(p, nA));                             // :234
    }
    if (xb && yb) return __ddiv_rn(two_p, nA);                                // :238
    if (xb || yb) return __ddiv_rn(p, nA);                                    // :242
    return 0.0;
}

// Rows of the states [lo, lo + cnt) only (cnt == S, lo == 0: the whole world); the arrays have
// stride cnt and hold GLOBAL neighbour indices.  p_slip_scalar is used when p_slip == nullptr.
// K slots per state: 5 (one per stencil position, the shape of the register-resident kernels) or 4
// (compact: a grid-world state never has more than 4 distinct successors / predecessors -- 4
// neighbours, or 3 + itself on an edge -- which cuts the bytes the streamed kernels move by 15-22 %).
__global__ void gridworld_tables_kernel(int n, int icy, int B, const double *__restrict__ p_slip,
                                        double p_slip_scalar, int lo, int cnt, int K,
                                        int32_t *succ_idx, double *succ_p,
                                        int32_t *pred_idx, double *pred_p) {
    constexpr int A = 4;
    const int S = cnt;                       // stride of the output arrays
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * S) return;
    const int b = (int)(gid / S), sl = (int)(gid % S);
    const double p = icy ? (p_slip ? p_slip[b] : p_slip_scalar) : 0.0;
    int32_t *si = succ_idx + (size_t)b * K * S, *pi = pred_idx + (size_t)b * K * S;
    double *sp = succ_p + (size_t)b * A * K * S, *pp = pred_p + (size_t)b * A * K * S;
    const int sg = lo + sl;                  // global state
    const int x = sg % n, y = sg / n;
    // candidate neighbours in ascending state order
    const int cand[5] = {sg - n, sg - 1, sg, sg + 1, sg + n};
    const int s = sl;                        // local slot column
    const bool ok[5] = {y > 0, x > 0, true, x < n - 1, y < n - 1};
    int js = 0, jp = 0;
    for (int c = 0; c < 5; ++c) {
        if (!ok[c]) continue;
        const int t = cand[c];
        double ps[A], pq[A];
        bool nzs = false, nzp = false;
        for (int a = 0; a < A; ++a) {
            ps[a] = world_prob(n, icy, p, sg, t, a);    // s -> t
            pq[a] = world_prob(n, icy, p, t, sg, a);    // t -> s
            nzs |= ps[a] != 0.0;
            nzp |= pq[a] != 0.0;
        }
        if (nzs && js < K) {
            si[(size_t)js * S + s] = t;
            for (int a = 0; a < A; ++a) sp[((size_t)a * K + js) * S + s] = ps[a];
            ++js;
        }
        if (nzp && jp < K) {
            pi[(size_t)jp * S + s] = t;
            for (int a = 0; a < A; ++a) pp[((size_t)a * K + jp) * S + s] = pq[a];
            ++jp;
        }
    }
    for (; js < K; ++js) {
        si[(size_t)js * S + s] = sg;
        for (int a = 0; a < A; ++a) sp[((size_t)a * K + js) * S + s] = 0.0;
    }
    for (; jp < K; ++jp) {
        pi[(size_t)jp * S + s] = sg;
        for (int a = 0; a < A; ++a) pp[((size_t)a * K + jp) * S + s] = 0.0;
    }
}

__global__ void gridworld_dense_kernel(int n, int icy, double p, double *P) {
    const int S = n * n;
    const long long total = (long long)S * S;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int sf = (int)(g / S), st = (int)(g % S);
        double v[4];
        const int d = abs(sf % n - st % n) + abs(sf / n - st / n);
        for (int a = 0; a < 4; ++a) v[a] = d <= 1 ? world_prob(n, icy, p, sf, st, a) : 0.0;
        double4 *out = reinterpret_cast<double4 *>(P + g * 4);
        // two 16-byte stores (double4 needs 32-byte alignment, which g*32 bytes has)
        *out = make_double4(v[0], v[1], v[2], v[3]);
    }
}

// ---------------------------------------------------------------------------
// dense feature products (maxent.py:244, :248)
// ---------------------------------------------------------------------------
// reward[s] = sum_f features[s][f] * theta[f]; one warp per state
__global__ void features_dot_kernel(const double *__restrict__ F, int S, int nF,
                                    const double *__restrict__ theta, double *reward) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= S) return;
    const double *row = F + (size_t)warp * nF;
    double acc = 0.0;
    for (int f = lane; f < nF; f += 32) acc = fma(row[f], theta[f], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) reward[warp] = acc;
}

// grad[f] = e_features[f] - sum_s features[s][f] * svf[s]; one thread per feature,
// CTA-strided over states with a fixed-order shared-memory combine (deterministic)
__global__ void __launch_bounds__(256)
features_grad_kernel(const double *__restrict__ F, int S, int nF, const double *__restrict__ svf,
                     const double *__restrict__ ef, double *grad) {
    // block = 32 features x 8 state-lanes
    const int fl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + fl;
    __shared__ double part[8][33];
    double acc = 0.0;
    if (f < nF)
        for (int s = sl; s < S; s += 8) acc = fma(F[(size_t)s * nF + f], svf[s], acc);
    part[sl][fl] = acc;
    __syncthreads();
    if (sl == 0 && f < nF) {
        double t = part[0][fl];
        for (int i = 1; i < 8; ++i) t += part[i][fl];
        grad[f] = ef[f] - t;
    }
}

}  // namespace irlb200

// ===========================================================================
// C ABI
// ===========================================================================
using namespace irlb200;

#define CHECK_LAUNCH(what)                                     \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return fail_cuda(e__, what);   \
    } while (0)

extern "C" int irlb200_dense_count(const double *P, int S, int A, int32_t *succ_cnt,
                                   int32_t *pred_cnt, int32_t *kmax, void *stream) {
    if (!P || !succ_cnt || !pred_cnt || !kmax || S <= 0 || A <= 0) return fail(IRLB200_EINVAL, "dense_count: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(pred_cnt, 0, sizeof(int32_t) * S, st);
    cudaMemsetAsync(kmax, 0, sizeof(int32_t) * 2, st);
    dense_rows_kernel<false><<<S, 256, 0, st>>>(P, S, A, 0, 0, succ_cnt, pred_cnt, kmax,
                                                nullptr, nullptr, nullptr, nullptr, nullptr);
    CHECK_LAUNCH("dense_rows_kernel<count>");
    max_count_kernel<<<(S + 255) / 256 > 1024 ? 1024 : (S + 255) / 256, 256, 0, st>>>(pred_cnt, S, kmax + 1);
    CHECK_LAUNCH("max_count_kernel");
    return IRLB200_OK;
}

extern "C" int irlb200_dense_fill(const double *P, int S, int A, int Ks, int Kp,
                                  int32_t *succ_idx, double *succ_p, int32_t *pred_idx,
                                  double *pred_p, int32_t *pred_cnt, void *stream) {
    if (!P || !succ_idx || !succ_p || !pred_idx || !pred_p || !pred_cnt || S <= 0 || A <= 0 || Ks <= 0 || Kp <= 0)
        return fail(IRLB200_EINVAL, "dense_fill: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *cursor = nullptr;
    if (int rc = workspace(2, sizeof(int32_t) * (size_t)S, (void **)&cursor, st)) return rc;
    cudaMemsetAsync(cursor, 0, sizeof(int32_t) * S, st);
    dense_rows_kernel<true><<<S, 256, 0, st>>>(P, S, A, Ks, Kp, nullptr, nullptr, nullptr,
                                               succ_idx, succ_p, pred_idx, pred_p, cursor);
    CHECK_LAUNCH("dense_rows_kernel<fill>");
    pred_sort_kernel<<<(S + 127) / 128, 128, 0, st>>>(S, A, Kp, pred_cnt, pred_idx, pred_p);
    CHECK_LAUNCH("pred_sort_kernel");
    return IRLB200_OK;
}

extern "C" int irlb200_gridworld_tables_k(int size, int icy, int B, const double *p_slip, int K,
                                          int32_t *succ_idx, double *succ_p,
                                          int32_t *pred_idx, double *pred_p, void *stream) {
    if (size <= 0 || B <= 0 || (icy && !p_slip) || !succ_idx || !succ_p || !pred_idx || !pred_p || (K != 4 && K != 5))
        return fail(IRLB200_EINVAL, "gridworld_tables: bad argument");
    if ((long long)size * size > 0x7fffffffLL) return fail(IRLB200_EINVAL, "gridworld_tables: S >= 2^31");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    const long long total = (long long)B * size * size;
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return fail(IRLB200_EINVAL, "gridworld_tables: batch too large");
    gridworld_tables_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        size, icy, B, p_slip, 0.0, 0, size * size, K, succ_idx, succ_p, pred_idx, pred_p);
    CHECK_LAUNCH("gridworld_tables_kernel");
    return IRLB200_OK;
}

extern "C" int irlb200_gridworld_tables(int size, int icy, int B, const double *p_slip,
                                        int32_t *succ_idx, double *succ_p,
                                        int32_t *pred_idx, double *pred_p, void *stream) {
    return irlb200_gridworld_tables_k(size, icy, B, p_slip, 5, succ_idx, succ_p, pred_idx, pred_p, stream);
}

extern "C" int irlb200_gridworld_tables_range_k(int size, int icy, double p_slip, int lo, int cnt, int K,
                                                int32_t *succ_idx, double *succ_p,
                                                int32_t *pred_idx, double *pred_p, void *stream) {
    if (size <= 0 || cnt <= 0 || lo < 0 || (long long)lo + cnt > (long long)size * size ||
        !succ_idx || !succ_p || !pred_idx || !pred_p || (K != 4 && K != 5))
        return fail(IRLB200_EINVAL, "gridworld_tables_range: bad argument");
    if ((long long)size * size > 0x7fffffffLL) return fail(IRLB200_EINVAL, "gridworld_tables_range: S >= 2^31");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    gridworld_tables_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        size, icy, 1, nullptr, p_slip, lo, cnt, K, succ_idx, succ_p, pred_idx, pred_p);
    CHECK_LAUNCH("gridworld_tables_kernel<range>");
    return IRLB200_OK;
}

extern "C" int irlb200_gridworld_tables_range(int size, int icy, double p_slip, int lo, int cnt,
                                              int32_t *succ_idx, double *succ_p,
                                              int32_t *pred_idx, double *pred_p, void *stream) {
    return irlb200_gridworld_tables_range_k(size, icy, p_slip, lo, cnt, 5, succ_idx, succ_p, pred_idx, pred_p, stream);
}

extern "C" int irlb200_gridworld_dense(int size, int icy, double p_slip, double *P, void *stream) {
    if (size <= 0 || !P) return fail(IRLB200_EINVAL, "gridworld_dense: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    gridworld_dense_kernel<<<sms * 8, 256, 0, (cudaStream_t)stream>>>(size, icy, p_slip, P);
    CHECK_LAUNCH("gridworld_dense_kernel");
    return IRLB200_OK;
}

extern "C" int irlb200_features_dot(const double *features, int S, int F, const double *theta,
                                    double *reward, void *stream) {
    if (!features || !theta || !reward || S <= 0 || F <= 0) return fail(IRLB200_EINVAL, "features_dot: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    const long long threads = (long long)S * 32;
    features_dot_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(features, S, F, theta, reward);
    CHECK_LAUNCH("features_dot_kernel");
    return IRLB200_OK;
}

extern "C" int irlb200_features_grad(const double *features, int S, int F, const double *svf,
                                     const double *e_features, double *grad, void *stream) {
    if (!features || !svf || !e_features || !grad || S <= 0 || F <= 0) return fail(IRLB200_EINVAL, "features_grad: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    features_grad_kernel<<<(F + 31) / 32, 256, 0, (cudaStream_t)stream>>>(features, S, F, svf, e_features, grad);
    CHECK_LAUNCH("features_grad_kernel");
    return IRLB200_OK;
}
