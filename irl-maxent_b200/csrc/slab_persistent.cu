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
#include "phases.cuh"

namespace irlb200 {

constexpr int kMaxRanks = 16;

struct SlabShared {                               // start of every rank's peer-mapped block
    unsigned long long flags[2][kMaxRanks];       // [seq parity][source rank]: (seq + 1) << 8 | votes
    unsigned long long slot[4];                   // local arrivals [19:0] + gt votes [39:20] + nan votes [59:40]
    unsigned long long release[4];                // (seq + 1) << 8 | decision bits {1: gt, 2: nan, 4: abort}
};
constexpr size_t kSlabHeaderBytes = 1024;

struct SlabPeers {
    SlabShared *shared[kMaxRanks];                // every rank's header (own entry = local pointer)
    double *lo_buf0, *lo_buf1;                    // iterate buffers of rank - 1 (null at the first rank)
    double *hi_buf0, *hi_buf1;                    // iterate buffers of rank + 1 (null at the last rank)
    int rank, world;
    int lo, hi, halo;                             // owned global range and ghost width
    long long timeout_ns;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

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
                                       int32_t *n_iter, int32_t *status, double timeout_s, void *stream) {
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
    int nb = 0;
    void *params[6];
    cudaError_t e;
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
        } else if (op == 1) {
            auto k = slab_succ_kernel<kOpSoftVI, 0, 0>;
            if (int rc = coop_blocks(k, cnt, threads, &nb)) return rc;
            e = cudaLaunchCooperativeKernel((const void *)k, dim3(nb), dim3(threads), params, 0, st);
        } else if (fast) {
            auto k = slab_succ_kernel<kOpVI, 4, 5>;
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
