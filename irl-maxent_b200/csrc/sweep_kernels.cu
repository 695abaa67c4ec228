// sweep_kernels.cu -- launchers, dispatch and the C ABI of the sweep kernels.
//
//   kernels_cta.cuh      *_cta_kernel: grid = B problems, one CTA each (iterate in shared memory);
//                        step_cta_kernel / step_warp_kernel: fused gradient step
//   kernels_tiled.cuh    svf_grid5_kernel / backward_grid5_kernel: stencil tiles in registers
//   kernels_cluster.cuh  one world per thread-block cluster (DSMEM), push exchange
//   here                 *_grid_kernel: one problem, cooperative persistent grid (iterate in L2/HBM)
#include <cooperative_groups.h>

#include <cstdio>
#include <mutex>

#include "host_util.h"
#include "phases.cuh"
#include "batch_args.cuh"
#include "kernels_cta.cuh"
#include "kernels_tiled.cuh"
#include "kernels_cluster.cuh"

namespace irlb200 {

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
// compact 4-slot grid-world tables (irlb200_gridworld_tables_k): streamed / cooperative-grid kernels
static inline bool is_compact_shape(int A, int K) { return A == 4 && K == 4; }

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
        if (int rc = workspace(0, (size_t)B * S * K * sizeof(double), (void **)&ws, st)) return rc;
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
        if (int rc = workspace(0, (size_t)B * S * bt.f.K * sizeof(double), (void **)&ws, st)) return rc;
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
    if (int rc = workspace(1, hdr + 2 * (size_t)S * sizeof(double), (void **)&base, st)) return rc;
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
        if (OP == kOpSoftVI && env_int("IRLB200_SOFTVI_OCC", 3) == 3) {
            auto k3 = succ_grid_kernel<OP, 4, 5, 0, 256, 3>;
            if (int rc = plan_grid(k3, a.S, 256, 1, &pl)) return rc;
            return launch_coop(k3, pl, a, w, n_iter, status, st);
        }
        auto k = succ_grid_kernel<OP, 4, 5, 0, 512, 1>;
        if (int rc = plan_grid(k, a.S, threads, 1, &pl)) return rc;
        return launch_coop(k, pl, a, w, n_iter, status, st);
    }
    if (is_compact_shape(a.A, a.K)) {
        // 4-slot grid-world tables: same two flavours, 15-22 % fewer bytes per streamed sweep
        auto kr = succ_grid_kernel<OP, 4, 4, 1, 256, 1>;
        GridPlan pr;
        if (int rc = plan_grid(kr, a.S, 256, 1, &pr)) return rc;
        if ((long long)pr.blocks * 256 >= a.S && !env_int("IRLB200_FORCE_STREAMED", 0))
            return launch_coop(kr, pr, a, w, n_iter, status, st);
        if (OP == kOpSoftVI && env_int("IRLB200_SOFTVI_OCC", 3) == 3) {
            // three CTAs of 256 threads per SM (75 registers): the exp / log work of one state then overlaps the
            // table loads of another -- 2048 x 2048: 166 -> 116 us per sweep (92 % of the HBM copy bandwidth)
            auto k3 = succ_grid_kernel<OP, 4, 4, 0, 256, 3>;
            if (int rc = plan_grid(k3, a.S, 256, 1, &pl)) return rc;
            return launch_coop(k3, pl, a, w, n_iter, status, st);
        }
        auto k = succ_grid_kernel<OP, 4, 4, 0, 512, 1>;
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
    const bool fast = is_fast_shape(a.A, a.K), compact = is_compact_shape(a.A, a.K);
    const int threads = env_int("IRLB200_GRID_THREADS", 256);
    if (fast) {
        auto kr = svf_grid_kernel<4, 5, 1, 256, 1>;
        GridPlan pr;
        if (int rc = plan_grid(kr, a.S, 256, 1, &pr)) return rc;
        if ((long long)pr.blocks * 256 >= a.S && !env_int("IRLB200_FORCE_STREAMED", 0)) {
            a.w_scratch = nullptr;
            return launch_coop(kr, pr, a, w, n_iter, status, st);
        }
    } else if (compact) {
        auto kr = svf_grid_kernel<4, 4, 1, 256, 1>;
        GridPlan pr;
        if (int rc = plan_grid(kr, a.S, 256, 1, &pr)) return rc;
        if ((long long)pr.blocks * 256 >= a.S && !env_int("IRLB200_FORCE_STREAMED", 0)) {
            a.w_scratch = nullptr;
            return launch_coop(kr, pr, a, w, n_iter, status, st);
        }
    }
    double *ws = nullptr;
    if (int rc = workspace(0, (size_t)a.S * a.K * sizeof(double), (void **)&ws, st)) return rc;
    a.w_scratch = ws;
    if (fast) {
        auto k = svf_grid_kernel<4, 5, 0, 512, 1>;
        if (int rc = plan_grid(k, a.S, threads, 1, &pl)) return rc;
        return launch_coop(k, pl, a, w, n_iter, status, st);
    }
    if (compact) {
        auto k = svf_grid_kernel<4, 4, 0, 512, 1>;
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
    if (mode == IRLB200_MODE_CLUSTER)
        return fail(IRLB200_ELIMIT, "thread-block-cluster kernels exist for irlb200_backward and irlb200_svf on grid-stencil "
                                    "tables (n <= 128, n % 4 == 0); this operator / table runs in CTA or cooperative-grid mode");
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
                              (mode == IRLB200_MODE_AUTO && cl_size > 0 && t->stencil_n > 64 &&
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
    return irlb200_svf_ordered(t, B, p_initial, p0_shared, terminal_mask, mask_shared, policy, eps, max_sweeps, svf,
                               e_features, ef_shared, grad, n_iter, status, mode, nullptr, stream);
}

extern "C" int irlb200_svf_ordered(const irlb200_tables *t, int B, const double *p_initial, int p0_shared,
                                   const uint8_t *terminal_mask, int mask_shared, const double *policy,
                                   double eps, int max_sweeps, double *svf, const double *e_features,
                                   int ef_shared, double *grad, int32_t *n_iter, int32_t *status, int mode,
                                   const int32_t *order, void *stream) {
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
    bt.order = (!want_cluster && mode == IRLB200_MODE_CTA) ? order : nullptr;   // one-CTA-per-problem launches only
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

// The outer loop of irl / irl_causal on the device (tiny worlds, identity features, built-in optimizer).
extern "C" int irlb200_irl_small(const irlb200_tables *t, int B, int causal, double *theta,
                                 const double *e_features, int ef_shared, const double *p_initial, int p0_shared,
                                 const uint8_t *terminal_mask, const double *phi, int mask_shared,
                                 int n_backward, double discount, double eps_lap, double eps_svf, int max_sweeps,
                                 int opt_kind, const double *lr, int lr_shared, int n_rates, double eps,
                                 int32_t *steps, int32_t *done, int32_t *last_counts, void *stream) {
    if (int rc = check_tables(t, true, true)) return rc;
    if (B <= 0 || !theta || !e_features || !p_initial || !terminal_mask || !lr || n_rates <= 0 || !steps || !done)
        return fail(IRLB200_EINVAL, "irl_small: bad argument");
    if (causal && !phi) return fail(IRLB200_EINVAL, "irl_small: causal needs phi");
    if (opt_kind != 0 && opt_kind != 1) return fail(IRLB200_EINVAL, "irl_small: optimizer kind must be 0 (Sga) or 1 (ExpSga)");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    if (t->S > 32 || !is_fast_shape(t->A, t->Ks) || !is_fast_shape(t->A, t->Kp))
        return fail(IRLB200_ELIMIT, "irl_small: needs S <= 32, A = 4 and 5-slot tables");
    StepBatch bt{};
    fill_succ(bt.s, t);
    bt.s.phi = phi; bt.s.term = terminal_mask; bt.s.discount = discount;
    bt.s.eps = eps_lap; bt.s.n_sweeps = n_backward; bt.s.max_sweeps = max_sweeps;
    fill_svf(bt.f, t);
    bt.f.p0 = p_initial; bt.f.term = terminal_mask; bt.f.eps = eps_svf; bt.f.max_sweeps = max_sweeps;
    bt.f.e_features = e_features;
    bt.succ_idx_stride = t->shared ? 0 : (size_t)t->Ks * t->S;
    bt.succ_p_stride = t->shared ? 0 : (size_t)t->A * t->Ks * t->S;
    bt.pred_idx_stride = t->shared ? 0 : (size_t)t->Kp * t->S;
    bt.pred_p_stride = t->shared ? 0 : (size_t)t->A * t->Kp * t->S;
    bt.phi_stride = mask_shared ? 0 : (size_t)t->S;
    bt.term_stride = mask_shared ? 0 : (size_t)t->S;
    bt.p0_stride = p0_shared ? 0 : (size_t)t->S;
    bt.ef_stride = ef_shared ? 0 : (size_t)t->S;
    IrlLoopArgs lp{};
    lp.theta = theta; lp.lr = lr; lp.lr_stride = lr_shared ? 0 : (size_t)n_rates; lp.n_rates = n_rates;
    lp.kind = opt_kind; lp.eps = eps; lp.steps = steps; lp.done = done; lp.last_counts = last_counts;
    cudaStream_t st = (cudaStream_t)stream;
    // few causal problems: latency bound -> four warps per problem (one per action); batches: one warp each
    if (causal && B <= env_int("IRLB200_CTA4_MAX_BATCH", 64)) irl_cta4_kernel<<<B, 128, 0, st>>>(bt, lp, B);
    else if (causal) irl_warp_kernel<true><<<(B + 3) / 4, 128, 0, st>>>(bt, lp, B);
    else irl_warp_kernel<false><<<(B + 3) / 4, 128, 0, st>>>(bt, lp, B);
    LAUNCH_CHECK("irl_warp_kernel");
    return IRLB200_OK;
}
