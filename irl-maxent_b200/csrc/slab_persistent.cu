// slab_persistent.cu -- slab mode with the halo exchange INSIDE the sweep kernel.
//
// One persistent cooperative kernel per GPU runs the whole fixed point of its row slab
// (same per-state arithmetic as everywhere: phases.cuh, streamed flavour).  There is no
// collective call and no host round trip per sweep:
//
//   * the iterate buffers of all ranks live in peer-mapped memory (cudaMalloc + CUDA IPC);
//     SlabTopo::store() writes a boundary-row value both locally and straight into the
//     neighbour's ghost row over NVLink (same global offset in the peer's buffer);
//   * the per-sweep fence is a two-level barrier: CTAs arrive on a local 64-bit counter
//     (arrivals + votes, as in GridTopo); the LAST arriver publishes (sequence, votes) to
//     every rank's flag table with system-scope stores, waits for all ranks' flags of the
//     same sequence, ORs the votes (the reference's stop rule needs the max over ALL
//     states) and releases the local CTAs with the global decision;
//   * flags are double-buffered by sequence parity, release words carry the sequence, so
//     nothing is ever reset while kernels run.
//
// Spin loops carry a wall-clock timeout (a missing peer aborts the launch instead of
// hanging the GPU): status IRLB200_ST_ABORTED.
#include <cstring>

#include "host_util.h"
#include "slab_common.cuh"

namespace irlb200 {

struct SlabTopo {
    double *buf0, *buf1;
    const SlabPeers *pe;                          // kernel parameter (constant bank)
    SlabShared *sh;
    unsigned seq;
    double *scratch;
    unsigned long long *s_word;
    int *flag;
    int lo_edge, hi_edge;                         // lo + halo, hi - halo
    double *plo0, *plo1, *phi0, *phi1;            // neighbours' iterate buffers (registers; null = none)
    bool dead;                                    // a peer timed out: only local barriers from now on

    __device__ __forceinline__ int rank() const { return blockIdx.x * blockDim.x + threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return gridDim.x * blockDim.x; }
    __device__ __forceinline__ double load(int b, int i) const { return ld_cg((b ? buf1 : buf0) + i); }
    // owned value: local store + push into the neighbours' ghost rows over NVLink
    __device__ __forceinline__ void store(int b, int i, double v) const {
        st_cg((b ? buf1 : buf0) + i, v);
        if (i < lo_edge && plo0) st_cg((b ? plo1 : plo0) + i, v);
        if (i >= hi_edge && phi0) st_cg((b ? phi1 : phi0) + i, v);
    }
    __device__ __forceinline__ void begin_phase() {
        if (threadIdx.x == 0) *flag = 0;
        __syncthreads();
    }

    // Two-level barrier.  Memory ordering: every CTA fences at GPU scope before its arrival
    // (cumulative over the CTA's stores, local and peer, via the preceding bar.sync); the last
    // arriver observes all arrivals, issues ONE system-scope fence, then publishes its flag to the
    // peers with relaxed system-scope stores (release pattern) and polls the peers' flags with
    // acquire loads at system scope -- a second full fence would also wait for the acknowledgements of
    // the flag stores it has just sent over NVLink -- before releasing the local CTAs at GPU scope.
    // System-scope fences are slow (~1.5 us): one on the critical path, none with a single rank.
    __device__ __forceinline__ unsigned barrier(unsigned long long inc) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned s4 = seq & 3u, par = seq & 1u;
            const unsigned long long want = (unsigned long long)(seq + 1u);
            __threadfence();
            const unsigned long long old = atomicAdd(&sh->slot[s4], inc);
            if ((old & 0xFFFFFull) == (unsigned long long)gridDim.x - 1ull) {
                const unsigned long long tot = old + inc;
                unsigned long long all = (((tot >> 20) & 0xFFFFFull) ? 1ull : 0ull) | ((tot >> 40) ? 2ull : 0ull);
                const int G = pe->world, me = pe->rank;
                if (G > 1 && !dead) {
                    const unsigned long long word = (want << 8) | all;
                    __threadfence_system();
                    for (int r = 0; r < G; ++r)
                        if (r != me) *(volatile unsigned long long *)&pe->shared[r]->flags[par][me] = word;
                    const unsigned long long t0 = globaltimer_ns();
                    bool aborted = false;
                    for (int r = 0; r < G && !aborted; ++r) {
                        if (r == me) continue;
                        const unsigned long long *f = &sh->flags[par][r];
                        for (unsigned spins = 0;; ++spins) {
                            const unsigned long long v = ld_acquire_sys(f);     // acquire: no full fence needed after
                            if ((v >> 8) == want) { all |= v & 0xFFull; break; }
                            if ((v >> 8) == (~0ull >> 8)) { aborted = true; break; }        // a peer aborted
                            if ((spins & 1023u) == 1023u && (long long)(globaltimer_ns() - t0) > pe->timeout_ns) {
                                aborted = true;
                                break;
                            }
                        }
                    }
                    if (aborted) {
                        all |= 4ull;
                        for (int r = 0; r < G; ++r)
                            if (r != me) *(volatile unsigned long long *)&pe->shared[r]->flags[par][me] = ~0ull;
                    }
                }
                sh->slot[(seq + 2u) & 3u] = 0ull;
                st_release_gpu(&sh->release[s4], (want << 8) | (all & 0xFFull));
            }
            unsigned long long rel;
            do { rel = ld_acquire_gpu(&sh->release[s4]); } while ((rel >> 8) != want);
            *s_word = rel & 0xFFull;
        }
        __syncthreads();
        const unsigned w = (unsigned)*s_word;
        if (w & 4u) dead = true;
        ++seq;
        return w;
    }
    __device__ __forceinline__ void sync() { (void)barrier(1ull); }

    __device__ __forceinline__ int vote(const Vote &v) {
        if (v.nan) *flag = 1;
        const int any = __syncthreads_or(v.gt ? 1 : 0);
        unsigned long long inc = 1ull;
        if (threadIdx.x == 0) {
            if (any) inc |= 1ull << 20;
            if (*flag) inc |= 1ull << 40;
        }
        const unsigned d = barrier(inc);
        if (d & 4u) return IRLB200_ST_ABORTED;
        if (d & 2u) return IRLB200_ST_NONFINITE;
        return (d & 1u) ? kContinue : IRLB200_ST_CONVERGED;
    }
    // only the backward pass rescales; it is not offered in slab mode (2S sweeps at S = 4.2 M)
    __device__ __forceinline__ double reduce_max(double x) { return block_max(x, scratch); }
};

__device__ __forceinline__ void carve_slab(SlabTopo &tp, unsigned char *base, const SlabPeers *pe, int S_total) {
    __shared__ double s_scratch[32];
    __shared__ unsigned long long s_word;
    __shared__ int s_flag;
    tp.sh = reinterpret_cast<SlabShared *>(base);
    tp.buf0 = reinterpret_cast<double *>(base + kSlabHeaderBytes);
    tp.buf1 = tp.buf0 + S_total;
    tp.pe = pe;
    tp.seq = 0;
    tp.scratch = s_scratch;
    tp.s_word = &s_word;
    tp.flag = &s_flag;
    tp.lo_edge = pe->lo + pe->halo;
    tp.hi_edge = pe->hi - pe->halo;
    tp.dead = false;
    tp.plo0 = pe->lo_buf0; tp.plo1 = pe->lo_buf1;
    tp.phi0 = pe->hi_buf0; tp.phi1 = pe->hi_buf1;
}

template <int OP, int A_T, int K_T>
__global__ void __launch_bounds__(256, 3)
    slab_succ_kernel(const SuccArgs a, const SlabPeers pe, unsigned char *base, int S_total, int32_t *n_iter,
                     int32_t *status) {
    SlabTopo tp;
    carve_slab(tp, base, &pe, S_total);
    succ_phase<SlabTopo, OP, A_T, K_T, 0>(tp, a, n_iter, status);
}

template <int A_T, int K_T>
__global__ void __launch_bounds__(256, 4)
    slab_svf_kernel(const SvfArgs a, const SlabPeers pe, unsigned char *base, int S_total, int32_t *n_iter,
                    int32_t *status) {
    SlabTopo tp;
    carve_slab(tp, base, &pe, S_total);
    svf_phase<SlabTopo, A_T, K_T, 0>(tp, a, n_iter, status);
}

// ---------------------------------------------------------------------------
// Boundary-first variant: the halo exchange overlaps the interior sweep.
//
// The first / last `halo` owned states are the only ones a neighbour needs.  One CTA per
// boundary row computes that row first, pushes it into the neighbour's ghost row, issues the
// (slow) system-scope fence and raises the neighbour's data flag -- all while the remaining
// CTAs sweep the interior.  The end-of-sweep barrier then needs no fence at all: the last
// local arriver publishes the rank's votes with relaxed stores, makes sure the neighbours'
// data flags of this sweep are in (they have long arrived) and waits for everybody's votes
// (one NVLink flight).  Same arithmetic, same stop rule, same results as slab_*_kernel.
// ---------------------------------------------------------------------------
template <int OP, int A_T, int K_T>
__global__ void __launch_bounds__(256, OP == 3 ? 4 : 3)
    slab_overlap_kernel(const OverlapArgs a, const SlabPeers pe, unsigned char *base, int32_t *n_iter, int32_t *status) {
    __shared__ unsigned long long s_word;
    __shared__ int s_nan;
    constexpr int QN = A_T > 0 ? A_T : kMaxDynA;
    SlabShared *sh = reinterpret_cast<SlabShared *>(base);
    double *buf0 = reinterpret_cast<double *>(base + kSlabHeaderBytes), *buf1 = buf0 + a.S_total;
    const int tid = threadIdx.x, cta = blockIdx.x, nthr = blockDim.x;
    const int G = pe.world, me = pe.rank, lo = a.lo, cnt = a.cnt, h = a.halo;
    const bool has_lo = pe.lo_buf0 != nullptr, has_hi = pe.hi_buf0 != nullptr;
    // roles: a group of NB CTAs per boundary row sweeps that row FIRST (about one state per thread, so the
    // row is complete after ~1 us); the last CTA of the group to finish fences and raises the neighbour's
    // flag; then every CTA takes its share of the interior
    const int NB = min(8, max(1, h / nthr));
    const bool low_role = has_lo && cta < NB;
    const bool high_role = has_hi && cta >= (has_lo ? NB : 0) && cta < (has_lo ? NB : 0) + NB;
    const int role_rank = low_role ? cta : cta - (has_lo ? NB : 0);          // index inside the boundary group
    const int i_begin = has_lo ? h : 0, i_end = has_hi ? cnt - h : cnt;       // interior range (local indices)
    const int A = A_T > 0 ? A_T : a.A, K = K_T > 0 ? K_T : a.K;
    unsigned seq = 0;
    bool dead = false;

    // the same two-level barrier as SlabTopo::barrier, with the cross-GPU part made fence-free:
    // `data_sync` = wait for the neighbours' data flags of this sequence number (boundary-first sweeps)
    auto barrier = [&](unsigned long long inc, bool fence_data, bool data_sync) -> unsigned {
        __syncthreads();
        if (tid == 0) {
            const unsigned s4 = seq & 3u, par = seq & 1u;
            const unsigned long long want = (unsigned long long)(seq + 1u);
            __threadfence();
            const unsigned long long old = atomicAdd(&sh->slot[s4], inc);
            if ((old & 0xFFFFFull) == (unsigned long long)gridDim.x - 1ull) {
                const unsigned long long tot = old + inc;
                unsigned long long all = (((tot >> 20) & 0xFFFFFull) ? 1ull : 0ull) | ((tot >> 40) ? 2ull : 0ull);
                if (G > 1 && !dead) {
                    if (fence_data) __threadfence_system();     // prologue / epilogue barriers publish data too
                    const unsigned long long word = (want << 8) | all;
                    for (int r = 0; r < G; ++r)
                        if (r != me) *(volatile unsigned long long *)&pe.shared[r]->flags[par][me] = word;
                    const unsigned long long t0 = globaltimer_ns();
                    bool aborted = false;
                    auto wait_for = [&](const unsigned long long *f, bool votes) {
                        for (unsigned spins = 0; !aborted; ++spins) {
                            const unsigned long long v = ld_acquire_sys(f);
                            const unsigned long long sq = votes ? (v >> 8) : v;
                            if (sq == want) { if (votes) all |= v & 0xFFull; return; }
                            if (v == ~0ull) { aborted = true; return; }
                            if ((spins & 1023u) == 1023u && (long long)(globaltimer_ns() - t0) > pe.timeout_ns) aborted = true;
                        }
                    };
                    if (data_sync) {
                        if (has_lo) wait_for(&sh->dflag[par][0], false);
                        if (has_hi) wait_for(&sh->dflag[par][1], false);
                    }
                    for (int r = 0; r < G; ++r)
                        if (r != me) wait_for(&sh->flags[par][r], true);
                    if (aborted) {
                        all |= 4ull;
                        for (int r = 0; r < G; ++r)
                            if (r != me) *(volatile unsigned long long *)&pe.shared[r]->flags[par][me] = ~0ull;
                        if (has_lo) *(volatile unsigned long long *)&pe.shared[me - 1]->dflag[par][1] = ~0ull;
                        if (has_hi) *(volatile unsigned long long *)&pe.shared[me + 1]->dflag[par][0] = ~0ull;
                    }
                }
                sh->slot[(seq + 2u) & 3u] = 0ull;
                st_release_gpu(&sh->release[s4], (want << 8) | (all & 0xFFull));
            }
            unsigned long long rel;
            do { rel = ld_acquire_gpu(&sh->release[s4]); } while ((rel >> 8) != want);
            s_word = rel & 0xFFull;
        }
        __syncthreads();
        const unsigned w = (unsigned)s_word;
        if (w & 4u) dead = true;
        ++seq;
        return w;
    };
    // value of an owned state: local store, plus the neighbour's ghost row for boundary states
    auto put = [&](int b, int i, double v) {
        const int g = lo + i;
        st_cg((b ? buf1 : buf0) + g, v);
        if (has_lo && i < h) st_cg((b ? pe.lo_buf1 : pe.lo_buf0) + g, v);
        if (has_hi && i >= cnt - h) st_cg((b ? pe.hi_buf1 : pe.hi_buf0) + g, v);
    };

    // ---- prologue: weights (forward pass), initial iterate ---------------------------------
    if (tid == 0) s_nan = 0;
    const int gtid = cta * nthr + tid, gthreads = gridDim.x * nthr;
    for (int i = gtid; i < cnt; i += gthreads) {
        if (OP == 3) {
            for (int j = 0; j < K; ++j) {
                const int pred = a.idx[(size_t)j * cnt + i];
                double acc = 0.0;
                for (int aa = 0; aa < A; ++aa)
                    acc = fma(__ldg(a.p + ((size_t)aa * K + j) * cnt + i), a.policy_in[(size_t)pred * A + aa], acc);
                a.w[(size_t)j * cnt + i] = a.term[pred] ? 0.0 : acc;
            }
        }
        put(0, i, OP == kOpSoftVI ? kNegHuge : 0.0);
    }
    (void)barrier(1ull, true, false);

    // ---- sweeps -----------------------------------------------------------------------------
    const int limit = a.max_sweeps > 0 ? a.max_sweeps : 0x7fffffff;
    int n = 0, st = IRLB200_ST_CONVERGED;
    for (;;) {
        const int b = n & 1;
        const double *x_in = b ? buf1 : buf0;
        bool gt = false, nan = false;
        auto sweep_range = [&](int i0, int i1, int first, int stride) {
            for (int i = i0 + first; i < i1; i += stride) {
                const double x = overlap_update<OP, A_T, K_T>(a, x_in, i, nullptr);
                const double diff = fabs(x - ld_cg(x_in + lo + i));
                gt |= diff > a.eps;
                nan |= diff != diff;
                put(b ^ 1, i, x);
            }
        };
        if (low_role || high_role) {
            if (low_role) sweep_range(0, h, role_rank * nthr + tid, NB * nthr);
            else sweep_range(cnt - h, cnt, role_rank * nthr + tid, NB * nthr);
            __syncthreads();
            if (tid == 0) {
                const unsigned par = seq & 1u;
                const int side = low_role ? 0 : 1;
                __threadfence();
                if (atomicAdd(&sh->bcount[par][side], 1u) == (unsigned)NB - 1u) {   // row complete
                    sh->bcount[par][side] = 0u;               // next use of this slot is two sweeps away
                    if (!dead) {
                        __threadfence_system();
                        unsigned long long *f = low_role ? &pe.shared[me - 1]->dflag[par][1]
                                                         : &pe.shared[me + 1]->dflag[par][0];
                        *(volatile unsigned long long *)f = (unsigned long long)(seq + 1u);
                    }
                }
            }
        }
        sweep_range(i_begin, i_end, cta * nthr + tid, gridDim.x * nthr);
        ++n;
        if (nan) s_nan = 1;
        const int any = __syncthreads_or(gt ? 1 : 0);
        unsigned long long inc = 1ull;
        if (tid == 0) {
            if (any) inc |= 1ull << 20;
            if (s_nan) inc |= 1ull << 40;
        }
        const unsigned d = barrier(inc, false, true);
        if (d & 4u) { st = IRLB200_ST_ABORTED; break; }
        if (d & 2u) { st = IRLB200_ST_NONFINITE; break; }
        if (!(d & 1u)) { st = IRLB200_ST_CONVERGED; break; }
        if (n >= limit) { st = IRLB200_ST_MAXSWEEPS; break; }
    }

    // ---- outputs ------------------------------------------------------------------------------
    {
        const double *x_new = (n & 1) ? buf1 : buf0, *x_old = ((n - 1) & 1) ? buf1 : buf0;
        for (int i = gtid; i < cnt; i += gthreads) {
            a.out[i] = ld_cg(x_new + lo + i);
            if (OP == kOpSoftVI && a.policy_out && n > 0) {
                double q[QN];
                const double x = overlap_update<OP, A_T, K_T>(a, x_old, i, q);
                for (int aa = 0; aa < A; ++aa) a.policy_out[(size_t)i * A + aa] = exp(q[aa] - x);     // maxent.py:341
            }
        }
    }
    if (gtid == 0) {
        if (n_iter) *n_iter = n;
        if (status) *status = st;
    }
    (void)barrier(1ull, false, false);     // nobody leaves while a peer may still push into its ghost rows
}

template <class Kern>
static int coop_blocks(Kern k, int cnt, int threads, int *blocks) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, 0);
    if (e != cudaSuccess) return fail_cuda(e, "occupancy");
    if (per_sm < 1) return fail(IRLB200_ELIMIT, "slab kernel does not fit on an SM");
    long long want = ((long long)cnt + threads - 1) / threads, cap = (long long)sms * per_sm;
    *blocks = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    return IRLB200_OK;
}

}  // namespace irlb200

using namespace irlb200;

// ---------------------------------------------------------------------------
// peer memory plumbing (CUDA IPC): plain cudaMalloc blocks that other ranks map
// ---------------------------------------------------------------------------
extern "C" int irlb200_peer_alloc(size_t bytes, void **ptr) {
    if (!ptr || bytes == 0) return fail(IRLB200_EINVAL, "peer_alloc: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(peer block)");
    e = cudaMemset(*ptr, 0, bytes);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemset(peer block)");
    return IRLB200_OK;
}
extern "C" int irlb200_peer_free(void *ptr) {
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "cudaFree(peer block)");
}
extern "C" int irlb200_ipc_export(void *ptr, unsigned char *handle64) {
    if (!ptr || !handle64) return fail(IRLB200_EINVAL, "ipc_export: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcGetMemHandle");
    memcpy(handle64, &h, 64);
    return IRLB200_OK;
}
extern "C" int irlb200_ipc_import(const unsigned char *handle64, void **ptr) {
    if (!ptr || !handle64) return fail(IRLB200_EINVAL, "ipc_import: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcOpenMemHandle");
    return IRLB200_OK;
}
extern "C" int irlb200_ipc_close(void *ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "cudaIpcCloseMemHandle");
}
extern "C" int irlb200_slab_reset(void *block, void *stream) {
    if (!block) return fail(IRLB200_EINVAL, "slab_reset: bad argument");
    cudaError_t e = cudaMemsetAsync(block, 0, kSlabHeaderBytes, (cudaStream_t)stream);
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "cudaMemsetAsync(slab header)");
}
extern "C" size_t irlb200_slab_block_bytes(int S_total) {
    return kSlabHeaderBytes + 2 * sizeof(double) * (size_t)S_total;
}

// One whole fixed point of the slab [lo, lo+cnt) in ONE persistent launch per rank.
//   blocks[r]  base pointer of rank r's peer block (own block for r == rank), world <= 16
//   op 1: soft-VI (p = succ_p, c0 = reward, c1 = phi, policy out [cnt][A], value out or NULL)
//   op 2: value iteration (p = succ_p, c0 = reward, value out [cnt])
//   op 3: forward pass (idx/p = predecessor tables, c0 = p_initial, policy_in [S_total][A] and
//         terminal_mask [S_total] global with ghost rows valid, w_scratch [K][cnt], out = svf [cnt])
extern "C" int irlb200_slab_persistent(int op, int rank, int world, void *const *blocks, int S_total, int lo,
                                       int cnt, int halo, int A, int K, const int32_t *idx, const double *p,
                                       const double *c0, const double *c1, const double *policy_in,
                                       const uint8_t *terminal_mask, double *w_scratch, double discount,
                                       double eps, int max_sweeps, int vi_mean, double *out, double *policy_out,
                                       int32_t *n_iter, int32_t *status, double timeout_s, int overlap,
                                       void *stream) {
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !blocks || cnt <= 0 || !idx || !p || !c0 || !out)
        return fail(IRLB200_EINVAL, "slab_persistent: bad argument");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    if (!(A == 4 && K == 5) && A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
    cudaStream_t st = (cudaStream_t)stream;
    SlabPeers pe{};
    for (int r = 0; r < world; ++r) pe.shared[r] = reinterpret_cast<SlabShared *>(blocks[r]);
    auto bufs = [&](int r, int which) {
        return reinterpret_cast<double *>(static_cast<unsigned char *>(blocks[r]) + kSlabHeaderBytes) + (size_t)which * S_total;
    };
    if (rank > 0) { pe.lo_buf0 = bufs(rank - 1, 0); pe.lo_buf1 = bufs(rank - 1, 1); }
    if (rank < world - 1) { pe.hi_buf0 = bufs(rank + 1, 0); pe.hi_buf1 = bufs(rank + 1, 1); }
    pe.rank = rank; pe.world = world; pe.lo = lo; pe.hi = lo + cnt; pe.halo = halo;
    pe.timeout_ns = (long long)((timeout_s > 0 ? timeout_s : 20.0) * 1e9);
    unsigned char *base = static_cast<unsigned char *>(blocks[rank]);
    // local barrier state starts from zero; the flag tables are zeroed by the CALLER on all ranks
    // before any rank launches (slab.py does it behind a process-group barrier)
    const int threads = 256;
    const bool fast = (A == 4 && K == 5);
    const bool compact = (A == 4 && K == 4);        // 4-slot grid-world tables: fewer bytes per streamed sweep
    int nb = 0;
    void *params[6];
    cudaError_t e;
    // overlap: 0 = never, 1 = where it pays (soft-VI / VI: measured 13 % faster at 2048^2 on 2 GPUs; the
    // forward sweep is too short for the boundary groups to get ahead of the interior), 2 = always
    if (((overlap == 1 && op != 3) || overlap >= 2) && world > 1 && cnt >= 2 * halo &&
        (op == 1 || op == 2 || op == 3)) {
        // boundary-first kernel: needs at least one interior CTA besides the (<= 2) boundary CTAs
        OverlapArgs oa{};
        oa.op = op; oa.lo = lo; oa.cnt = cnt; oa.S_total = S_total; oa.halo = halo; oa.A = A; oa.K = K;
        oa.idx = idx; oa.p = p; oa.c0 = c0; oa.c1 = c1; oa.policy_in = policy_in; oa.term = terminal_mask;
        oa.w = w_scratch; oa.discount = discount; oa.eps = eps; oa.max_sweeps = max_sweeps; oa.vi_mean = vi_mean;
        oa.out = out; oa.policy_out = policy_out;
        if (op == 3 && (!policy_in || !terminal_mask || !w_scratch)) return fail(IRLB200_EINVAL, "slab_persistent: forward pass inputs");
        if (op == 1 && (!c1 || !policy_out)) return fail(IRLB200_EINVAL, "slab_persistent: soft-VI inputs");
        void *op_params[5] = {&oa, &pe, &base, &n_iter, &status};
        const void *k = nullptr;
        if (op == 3) k = fast ? (const void *)slab_overlap_kernel<3, 4, 5> : compact ? (const void *)slab_overlap_kernel<3, 4, 4> : (const void *)slab_overlap_kernel<3, 0, 0>;
        else if (op == 1) k = fast ? (const void *)slab_overlap_kernel<kOpSoftVI, 4, 5> : compact ? (const void *)slab_overlap_kernel<kOpSoftVI, 4, 4> : (const void *)slab_overlap_kernel<kOpSoftVI, 0, 0>;
        else k = fast ? (const void *)slab_overlap_kernel<kOpVI, 4, 5> : compact ? (const void *)slab_overlap_kernel<kOpVI, 4, 4> : (const void *)slab_overlap_kernel<kOpVI, 0, 0>;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, 0);
        if (e != cudaSuccess) return fail_cuda(e, "occupancy(slab overlap)");
        long long want = ((long long)cnt + threads - 1) / threads, cap = (long long)sms * per_sm;
        nb = (int)(want < cap ? want : cap);
        if (nb >= 17) {      // two boundary groups of up to 8 CTAs and at least one more
            e = cudaLaunchCooperativeKernel(k, dim3(nb), dim3(threads), op_params, 0, st);
            if (e != cudaSuccess) return fail_cuda(e, "cudaLaunchCooperativeKernel(slab overlap)");
            return IRLB200_OK;
        }
    }
    if (op == 3) {
        if (!policy_in || !terminal_mask || !w_scratch) return fail(IRLB200_EINVAL, "slab_persistent: forward pass inputs");
        SvfArgs a{};
        a.S = cnt; a.A = A; a.K = K; a.idx = idx; a.p = p; a.p0 = c0; a.term = terminal_mask; a.policy = policy_in;
        a.w_scratch = w_scratch; a.s_off = lo; a.eps = eps; a.max_sweeps = max_sweeps; a.svf = out;
        params[0] = &a; params[1] = &pe; params[2] = &base; params[3] = &S_total; params[4] = &n_iter; params[5] = &status;
        if (fast) {
            auto k = slab_svf_kernel<4, 5>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        } else if (compact) {
            auto k = slab_svf_kernel<4, 4>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        } else {
            auto k = slab_svf_kernel<0, 0>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        }
    } else if (op == 1 || op == 2) {
        if (op == 1 && (!c1 || !policy_out)) return fail(IRLB200_EINVAL, "slab_persistent: soft-VI inputs");
        SuccArgs a{};
        a.S = cnt; a.A = A; a.K = K; a.idx = idx; a.p = p; a.reward = c0; a.phi = c1; a.discount = discount;
        a.eps = eps; a.max_sweeps = max_sweeps; a.vi_mean = vi_mean; a.s_off = lo;
        a.policy = op == 1 ? policy_out : nullptr; a.value = out;
        params[0] = &a; params[1] = &pe; params[2] = &base; params[3] = &S_total; params[4] = &n_iter; params[5] = &status;
        if (op == 1 && fast) {
            auto k = slab_succ_kernel<kOpSoftVI, 4, 5>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        } else if (op == 1 && compact) {
            auto k = slab_succ_kernel<kOpSoftVI, 4, 4>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        } else if (op == 1) {
            auto k = slab_succ_kernel<kOpSoftVI, 0, 0>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        } else if (fast) {
            auto k = slab_succ_kernel<kOpVI, 4, 5>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        } else if (compact) {
            auto k = slab_succ_kernel<kOpVI, 4, 4>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        } else {
            auto k = slab_succ_kernel<kOpVI, 0, 0>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        }
    } else {
        return fail(IRLB200_EINVAL, "slab_persistent: unknown op");
    }
    if (e != cudaSuccess) return fail_cuda(e, "cudaLaunchCooperativeKernel(slab)");
    return IRLB200_OK;
}
