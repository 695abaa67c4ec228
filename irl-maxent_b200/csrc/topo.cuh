// topo.cuh -- where the iterate vector lives and how one sweep is fenced from
// the next.  The per-state arithmetic (phases.cuh) is written once against
// this small interface:
//
//   rank(), nthreads()           this thread's index / the number of threads that
//                                share the problem
//   load(buf, i) / store(...)    read / write entry i of iterate buffer 0 or 1
//   sync()                       every store of the sweep is visible to every thread
//   vote(v)                      sync() + the reference's stop rule; returns
//                                kContinue or an IRLB200_ST_* status
//   reduce_max(x)                sync() + max of x over all threads
//
//   CtaTopo   one CTA owns the problem; iterate in shared memory; one
//             bar.red per sweep does both the fence and the vote.
//   GridTopo  a cooperative persistent grid owns the problem; iterate in
//             global memory (L2-resident up to ~10^6 states, HBM beyond);
//             one 64-bit atomic per CTA per sweep carries arrival + votes.
#pragma once
#include "common.cuh"

namespace irlb200 {

// ---------------------------------------------------------------------------
struct CtaTopo {
    double *buf0, *buf1; // shared memory, S doubles each (selected, never indexed: stays in registers)
    double *scratch;     // shared memory, 32 doubles
    int *flag;           // shared memory, 2 ints: sticky "a NaN diff was seen", by vote parity
    unsigned vseq;       // vote counter (same on every thread)

    __device__ __forceinline__ int rank() const { return threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    __device__ __forceinline__ double load(int b, int i) const { return (b ? buf1 : buf0)[i]; }
    __device__ __forceinline__ void store(int b, int i, double v) const { (b ? buf1 : buf0)[i] = v; }
    __device__ __forceinline__ void sync() { __syncthreads(); }

    __device__ __forceinline__ void begin_phase() {
        __syncthreads();
        if (threadIdx.x == 0) { flag[0] = 0; flag[1] = 0; }
        vseq = 0;
        __syncthreads();
    }
    // One bar.red per sweep is both the fence between sweeps and the stop rule.
    // The NaN flag is double-buffered by vote parity: a thread that is already
    // in sweep t+1 writes flag[(t+1)&1] while a slower one still reads flag[t&1].
    __device__ __forceinline__ int vote(const Vote &v) {
        int *f = flag + (vseq & 1u);
        ++vseq;
        if (v.nan) *f = 1;
        const int any = __syncthreads_or(v.gt ? 1 : 0);
        if (*f) return IRLB200_ST_NONFINITE;
        return any ? kContinue : IRLB200_ST_CONVERGED;
    }
    __device__ __forceinline__ double reduce_max(double x) { return block_max(x, scratch); }
};

// ---------------------------------------------------------------------------
// Global state of one cooperative launch.  Zeroed by the host before launch.
struct GridSyncState {
    unsigned long long slot[4];     // [19:0] arrivals, [39:20] "gt" votes, [59:40] "nan" votes
    unsigned long long maxbits[4];  // bit pattern of a non-negative double
};

struct GridTopo {
    double *buf0, *buf1;        // global memory, S doubles each
    GridSyncState *gs;
    unsigned seq;               // barrier sequence number (same on every thread)
    double *scratch;            // shared, 32 doubles
    unsigned long long *s_word; // shared, 1 word: what thread 0 saw at the barrier
    int *flag;                  // shared, sticky NaN flag of the CTA

    __device__ __forceinline__ int rank() const { return blockIdx.x * blockDim.x + threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return gridDim.x * blockDim.x; }
    __device__ __forceinline__ double load(int b, int i) const { return ld_cg((b ? buf1 : buf0) + i); }
    __device__ __forceinline__ void store(int b, int i, double v) const { st_cg((b ? buf1 : buf0) + i, v); }

    __device__ __forceinline__ void begin_phase() {
        if (threadIdx.x == 0) *flag = 0;
        __syncthreads();
    }

    // arrive with `inc`, wait for everybody, broadcast the final word to the CTA
    __device__ __forceinline__ unsigned long long barrier(unsigned long long inc) {
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long *s = &gs->slot[seq & 3u];
            __threadfence();
            atomicAdd(s, inc);
            unsigned long long c;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(s) : "memory");
            } while ((c & 0xFFFFFull) != (unsigned long long)gridDim.x);
            if (blockIdx.x == 0) {              // recycle the slots used two barriers from now
                gs->slot[(seq + 2u) & 3u] = 0ull;
                gs->maxbits[(seq + 2u) & 3u] = 0ull;
            }
            *s_word = c;
        }
        __syncthreads();
        const unsigned long long w = *s_word;
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
        const unsigned long long w = barrier(inc);
        if (w >> 40) return IRLB200_ST_NONFINITE;
        return ((w >> 20) & 0xFFFFFull) ? kContinue : IRLB200_ST_CONVERGED;
    }

    // x must be >= 0 or NaN (partition values): its bit pattern orders like the value
    __device__ __forceinline__ double reduce_max(double x) {
        double m = block_max(x, scratch);
        if (threadIdx.x == 0)
            atomicMax(&gs->maxbits[seq & 3u], (unsigned long long)__double_as_longlong(m));
        const unsigned slot = seq & 3u;
        (void)barrier(1ull);
        return __longlong_as_double((long long)ld_relaxed_u64(&gs->maxbits[slot]));
    }
};

}  // namespace irlb200
