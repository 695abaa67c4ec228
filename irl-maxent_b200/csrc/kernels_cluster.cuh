// kernels_cluster.cuh -- one world per thread-block cluster: barrier.cluster forward pass, push (st.async + mbarrier) forward and backward passes.
#pragma once
#include <cooperative_groups.h>
#include "kernels_tiled.cuh"

namespace irlb200 {

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
    stencil_tile_update<TY, TX, true>(w, p0r, cur, up, dn, lf, rt, x);
#pragma unroll
    for (int c = 0; c < TY * TX; ++c) *reinterpret_cast<double *>(smem + own + 8 * c + OFF_W) = x[c];
}

// cluster-wide OR of a per-thread predicate: stamp, barrier.cluster, compare.  `word` points at this
// CTA's vote words [2]; every CTA's copy is written by every voting warp (one lane per target CTA).
// Stamps only grow, and a CTA goes on to write a LARGER stamp into the same word only after it has seen
// this vote come out true (a false vote ends the loop or goes to the other word first) -- so a thread that
// reads the word late and finds a newer stamp knows this vote was true: compare with >=, not ==.
__device__ __forceinline__ bool cluster_any(cgx::cluster_group &cl, int *word, int stamp, bool pred, int ncta) {
    int *slot = word + (stamp & 1);
    const unsigned any = __ballot_sync(0xffffffffu, pred);
    const int lane = threadIdx.x & 31;
    if (any && lane < ncta) *cl.map_shared_rank(slot, lane) = stamp;
    cl.sync();
    return *(volatile int *)slot >= stamp;
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
            svf_stencil_weights<A, K>(a, S, s, n, live, w[c]);
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
    stencil_tile_update<TY, TX, true>(w, p0r, cur, up, dn, lf, rt, x);
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
            svf_stencil_weights<A, K>(a, S, s, n, live, w[c]);
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
    stencil_tile_update<TY, TX, false>(w, cur, cur, up, dn, lf, rt, x);
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
            const double z0 = backward_stencil_weights<A, K>(a, S, s, n, live, w[c], max_abs_r);
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


}  // namespace irlb200
