// slab_flow.cu -- slab mode without a per-sweep barrier ("dataflow" sweeps).
//
// slab_persistent.cu fences every sweep with a full barrier: all CTAs of the GPU arrive on one
// counter, the last one exchanges (sequence, votes) with EVERY rank and releases the others
// (2.7 us locally + an all-to-all NVLink flag round per sweep, against 6 us of HBM time per sweep
// of a 2048 x 2048 world on 8 GPUs).  A grid world's sweep only couples a state to the states one
// grid row away, so nothing in the arithmetic needs that:
//
//   * every persistent CTA owns a FIXED contiguous range of the slab's states and publishes, after
//     each sweep, one progress word.  Before sweep k it waits only for the few CTAs whose ranges lie
//     within one grid row of its own to have finished sweep k-1 -- which is both the
//     read-after-write condition (their rows are my halo) and the write-after-read condition (they
//     have finished reading the buffer I am about to overwrite).  CTAs of one SM drift apart by a
//     fraction of a sweep, so one CTA's wait is hidden behind another's loads;
//   * across GPUs the boundary rows travel through "LL" mailboxes (slab_common.cuh): the CTAs that
//     own the first / last grid row of a slab store every new value straight into the neighbouring
//     GPU's mailbox as two 8-byte words that each carry half the value and the iterate number.  The
//     consumer spins on the words it needs until both tags match: data and arrival in ONE one-way
//     NVLink flight -- no system-scope fence, no flag, no collective call on the sweep path.  Those
//     rows get their own small CTAs (about one state per thread), so the flight runs while the
//     interior CTAs are still sweeping;
//   * the stop rule (`while delta > eps`, maxent.py:108,326; solver.py:40) needs the maximum over
//     ALL states, but not immediately: every CTA records one vote bit per sweep, and only every
//     `chunk` sweeps the masks are OR-ed over the GPU and exchanged between the ranks behind one
//     full barrier.  To return exactly the iterate and the sweep count of the reference, the
//     iterate at the start of each chunk is snapshotted on the fly (one extra 8-byte store per state
//     in the first sweep of a chunk); when the masks show that the loop ended at sweep i of the
//     chunk, the snapshot is restored and sweeps 0..i are replayed (deterministic arithmetic:
//     bitwise the same values).  Cost: < 2 chunks of extra sweeps per fixed point.
//
// Spin loops carry a wall-clock timeout and watch an abort word, so a missing peer ends the launch
// with IRLB200_ST_ABORTED instead of hanging the GPU.  Same per-state arithmetic (slab_common.cuh)
// and therefore bitwise the same results as every other kernel of the library.
#include <cstdlib>

#include "host_util.h"
#include "slab_common.cuh"

namespace irlb200 {

struct FlowShared {                               // at byte kFlowOffset of every rank's peer-mapped header
    unsigned long long abort;                     // != 0: somebody timed out
    unsigned long long pad[7];
    unsigned long long vgt[2], vnan[2];           // vote masks of the current chunk (bit i = sweep i), by barrier parity
    unsigned long long rel_gt[4], rel_nan[4];     // decision masks broadcast with the barrier release
    unsigned long long xgt[2][kMaxRanks], xnan[2][kMaxRanks];   // masks received from the other ranks
};
constexpr size_t kFlowOffset = 1024;
static_assert(sizeof(SlabShared) <= kFlowOffset, "SlabShared grew into the flow header");
static_assert(kFlowOffset + sizeof(FlowShared) <= kSlabHeaderBytes, "flow header does not fit");
constexpr int kProgressStride = 4;                // progress words 32 bytes apart (one sector each)

// block of a rank: [header | iterate 0 [S_total] | iterate 1 [S_total] | mailboxes [2 sides][2 parities][halo]]
__host__ __device__ inline size_t flow_mail_offset(int S_total) {
    return (kSlabHeaderBytes + 2 * sizeof(double) * (size_t)S_total + 255) & ~(size_t)255;
}

__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
    *(volatile unsigned long long *)p = v;
}

// U > 1 (forward pass, compile-time K): a thread works on U of its states at a time -- all their table
// rows are requested first, then all gathers, so a sweep over ~3.5 states per thread (2048 x 2048 on
// 8 GPUs) is one or two rounds of dependent L2 accesses instead of 3.5.
// CODED (soft-VI / VI): the successor probabilities are read as one byte per entry and decoded through a
// 256-entry dictionary in shared memory -- bitwise the same doubles, 16 instead of 128 bytes per state and sweep
// for a 4-action, 4-slot table (2048 x 2048: 168 -> 56 B per state, which turns the sweep from HBM bound into
// FP64 bound).
template <int OP, int A_T, int K_T, int U, int MINB, bool CODED = false>
__global__ void __launch_bounds__(256, MINB)
    slab_flow_kernel(const OverlapArgs a, const SlabPeers pe, unsigned char *base, unsigned long long *progress,
                     double *snap, const int chunk, const int edge, const int res_cap, int32_t *n_iter,
                     int32_t *status) {
    extern __shared__ __align__(16) unsigned char res_raw[];   // forward pass: table rows of the CTA's first states
    __shared__ unsigned long long s_gt, s_nanmask;
    __shared__ int s_nan, s_dead;
    __shared__ double s_dict[CODED ? 256 : 1];
    constexpr int QN = A_T > 0 ? A_T : kMaxDynA;
    if (CODED) {
        for (int t = threadIdx.x; t < 256; t += blockDim.x) s_dict[t] = a.dict[t];
        __syncthreads();
    }
    SlabShared *sh = reinterpret_cast<SlabShared *>(base);
    FlowShared *fs = reinterpret_cast<FlowShared *>(base + kFlowOffset);
    double *buf0 = reinterpret_cast<double *>(base + kSlabHeaderBytes), *buf1 = buf0 + a.S_total;
    const int tid = threadIdx.x, cta = blockIdx.x, nthr = blockDim.x, nb = gridDim.x;
    const int G = pe.world, me = pe.rank, lo = a.lo, cnt = a.cnt, h = a.halo, hi = a.lo + a.cnt;
    const bool has_lo = me > 0, has_hi = me < G - 1;
    const int A = A_T > 0 ? A_T : a.A, K = K_T > 0 ? K_T : a.K;
    // mailboxes: mine (what the neighbours send me) and theirs (where my boundary rows go)
    const size_t mo = flow_mail_offset(a.S_total);
    LLSlot *my_mail = reinterpret_cast<LLSlot *>(base + mo);                       // [side 0: from rank-1 | side 1: from rank+1][parity][h]
    LLSlot *to_lo = has_lo ? reinterpret_cast<LLSlot *>(reinterpret_cast<unsigned char *>(pe.shared[me - 1]) + mo) + (size_t)2 * h : nullptr;
    LLSlot *to_hi = has_hi ? reinterpret_cast<LLSlot *>(reinterpret_cast<unsigned char *>(pe.shared[me + 1]) + mo) : nullptr;

    // ---- fixed ownership and the dependence set ----------------------------------------------
    // Three zones of contiguous, near-equal ranges: the slab's first grid row (when rank-1 exists) is split
    // over `edge` CTAs, its last row (when rank+1 exists) over another `edge`, the interior over the rest.
    // About one state per thread on the edge rows: their CTAs finish a sweep in a fraction of the time the
    // interior CTAs need, so the NVLink flight of the boundary row runs while the interior is still being
    // swept instead of adding to every sweep.  edge == 0: uniform split.
    const int eL = has_lo ? edge : 0, eH = has_hi ? edge : 0, mid = nb - eL - eH;
    const int z0 = eL ? h : 0, z1 = eH ? cnt - h : cnt;
    auto ceil_div = [](long long x, long long y) { return (int)((x + y - 1) / y); };
    auto beg_of = [&](int c) -> int {
        if (c < eL) return (int)((long long)z0 * c / eL);
        if (c >= nb - eH) return z1 + (int)((long long)(cnt - z1) * (c - (nb - eH)) / (eH ? eH : 1));
        return z0 + (int)((long long)(z1 - z0) * (c - eL) / mid);
    };
    auto cta_of = [&](int i) -> int {                         // owner of local state i
        if (i < z0) return ceil_div(((long long)i + 1) * eL, z0) - 1;
        if (i >= z1) return nb - eH + ceil_div(((long long)(i - z1) + 1) * eH, cnt - z1) - 1;
        return eL + ceil_div(((long long)(i - z0) + 1) * mid, z1 - z0) - 1;
    };
    const int beg = beg_of(cta), end = beg_of(cta + 1);
    const int d_lo = cta_of(max(beg - h, 0)), d_hi = cta_of(min(end + h, cnt) - 1);
    // within one grid row of the neighbouring slab: reads its boundary row from the mailbox
    const bool near_lo = has_lo && beg < h, near_hi = has_hi && end > cnt - h;
    // Forward pass: a CTA's range is FIXED, so the rows of its first `nres` states (predecessor indices, merged
    // weights, p_initial: 4 K + 8 K + 8 bytes per state) are copied into shared memory once and every sweep reads
    // them from there -- at 8 GPUs a 2048 x 2048 slab is ~900 states per CTA and fits entirely (198 kB per SM),
    // leaving only the iterate on the L2 path.  The launcher asks for it only when EVERY CTA's range fits
    // (res_cap >= the longest range); otherwise res_cap = 0 and all rows are streamed.
    constexpr int KR = K_T > 0 ? K_T : 1;
    const int nres = (OP == 3 && K_T > 0) ? min(end - beg, res_cap) : 0;
    int *s_idx = reinterpret_cast<int *>(res_raw);                                            // [K][nres]
    double *s_w = reinterpret_cast<double *>(res_raw + (((size_t)KR * nres * 4 + 15) & ~(size_t)15));   // [K][nres]
    double *s_p0 = s_w + (size_t)KR * nres;                                                   // [nres]

    unsigned seq = 0;                  // full-barrier sequence number
    bool dead = false;

    auto raise_abort = [&]() {
        st_volatile_u64(&fs->abort, 1ull);
        for (int r = 0; r < G; ++r)
            if (r != me)
                st_volatile_u64(&reinterpret_cast<FlowShared *>(reinterpret_cast<unsigned char *>(pe.shared[r]) + kFlowOffset)->abort, 1ull);
    };
    // slow path of every spin loop (every 1024 polls): false = give up
    auto keep_waiting = [&](unsigned long long &t0) -> bool {
        if (*(volatile unsigned long long *)&fs->abort) return false;
        const unsigned long long now = globaltimer_ns();
        if (!t0) t0 = now;
        else if ((long long)(now - t0) > pe.timeout_ns) { raise_abort(); return false; }
        return true;
    };
    // wait until *p >= want (monotonic words); false = aborted / timed out
    auto spin_ge = [&](const unsigned long long *p, unsigned long long want, bool sys) -> bool {
        unsigned long long t0 = 0ull;
        for (unsigned spins = 0;; ++spins) {
            const unsigned long long v = sys ? ld_acquire_sys(p) : ld_acquire_gpu(p);
            if (v >= want) return true;
            if ((spins & 1023u) == 1023u && !keep_waiting(t0)) return false;
        }
    };
    // iterate number `it` (buffer it & 1, mailbox tag it + 1) at global index g.
    // Mailbox protocol: the slot of iterate `it` is reused by iterate it + 2, and the reader insists on the exact
    // tag.  That is safe because the coupling across a slab boundary is one-to-one and symmetric in every grid
    // world (my state s reads the neighbour's s +- halo and nothing else over there, and that state reads s):
    // the neighbour cannot produce iterate it + 2 of the slot before it has received my iterate it + 1 of s,
    // which the thread that owns s sends only after it has read the slot.  Tables that couple slabs in any
    // other way must use irlb200_slab_persistent (slab.PeerSlabGrid(flow=False)).
    auto load_x = [&](const double *x_in, unsigned it, int g) -> double {
        if (g >= lo && g < hi) return ld_cg(x_in + g);
        const LLSlot *slot = g < lo ? my_mail + (size_t)(it & 1u) * h + (g - (lo - h))
                                    : my_mail + (size_t)(2u + (it & 1u)) * h + (g - hi);
        double v = 0.0;
        unsigned long long t0 = 0ull;
        for (unsigned spins = 0; !ll_try_read(slot, it + 1u, v); ++spins)
            if ((spins & 1023u) == 1023u && !keep_waiting(t0)) { s_dead = 1; break; }
        return v;
    };
    // iterate number `it` of an owned state: local store, plus the neighbour's mailbox for boundary rows
    auto put = [&](unsigned it, int i, double v) {
        st_cg(((it & 1u) ? buf1 : buf0) + lo + i, v);
        if (has_lo && i < h) ll_write(to_lo + (size_t)(it & 1u) * h + i, v, it + 1u);
        if (has_hi && i >= cnt - h) ll_write(to_hi + (size_t)(it & 1u) * h + (i - (cnt - h)), v, it + 1u);
    };
    // thread 0, after a bar.sync that covers the CTA's stores: "my rows hold iterate number v - 1"
    auto publish = [&](unsigned long long v) {
        __threadfence();
        st_volatile_u64(progress + (size_t)cta * kProgressStride, v);
    };

    // Full barrier over all CTAs of all ranks (once per chunk, not per sweep): local arrival counter, the
    // last arriver ORs the chunk's vote masks, exchanges them with every rank and releases the others.
    // Returns false when the launch is aborted.
    auto full_barrier = [&](unsigned long long my_gt, unsigned long long my_nan, unsigned long long &all_gt,
                            unsigned long long &all_nan) -> bool {
        __syncthreads();
        if (tid == 0) {
            const unsigned s4 = seq & 3u, par = seq & 1u;
            const unsigned long long want = (unsigned long long)(seq + 1u);
            if (my_gt) atomicOr(&fs->vgt[par], my_gt);
            if (my_nan) atomicOr(&fs->vnan[par], my_nan);
            __threadfence();
            const unsigned long long old = atomicAdd(&sh->slot[s4], 1ull);
            if (old == (unsigned long long)nb - 1ull) {
                __threadfence();
                unsigned long long mg = *(volatile unsigned long long *)&fs->vgt[par];
                unsigned long long mn = *(volatile unsigned long long *)&fs->vnan[par];
                fs->vgt[par] = 0ull;                         // next use: two barriers from now
                fs->vnan[par] = 0ull;
                bool ok = !dead && !s_dead;
                if (G > 1 && ok) {
                    for (int r = 0; r < G; ++r) {
                        if (r == me) continue;
                        FlowShared *fr = reinterpret_cast<FlowShared *>(reinterpret_cast<unsigned char *>(pe.shared[r]) + kFlowOffset);
                        st_volatile_u64(&fr->xgt[par][me], mg);
                        st_volatile_u64(&fr->xnan[par][me], mn);
                    }
                    __threadfence_system();                  // masks before the flags
                    for (int r = 0; r < G; ++r)
                        if (r != me) st_volatile_u64(&pe.shared[r]->flags[par][me], want);
                    for (int r = 0; r < G && ok; ++r) {
                        if (r == me) continue;
                        ok = spin_ge(&sh->flags[par][r], want, true);
                        if (ok) {
                            mg |= *(volatile unsigned long long *)&fs->xgt[par][r];
                            mn |= *(volatile unsigned long long *)&fs->xnan[par][r];
                        }
                    }
                }
                sh->slot[(seq + 2u) & 3u] = 0ull;
                fs->rel_gt[s4] = mg;
                fs->rel_nan[s4] = mn;
                st_release_gpu(&sh->release[s4], (want << 8) | (ok ? 0ull : 4ull));
            }
            unsigned long long rel, t0 = 0ull;
            for (unsigned spins = 0;; ++spins) {
                rel = ld_acquire_gpu(&sh->release[s4]);
                if ((rel >> 8) == want) break;
                if ((spins & 1023u) == 1023u) {                 // the last arriver itself gives up after timeout_ns
                    const unsigned long long now = globaltimer_ns();
                    if (!t0) t0 = now;
                    else if ((long long)(now - t0) > 3 * pe.timeout_ns) { rel = 4ull; break; }
                }
            }
            s_gt = *(volatile unsigned long long *)&fs->rel_gt[s4];
            s_nanmask = *(volatile unsigned long long *)&fs->rel_nan[s4];
            if (rel & 4ull) s_dead = 1;
        }
        __syncthreads();
        all_gt = s_gt;
        all_nan = s_nanmask;
        ++seq;
        if (s_dead) dead = true;
        return !dead;
    };

    // ---- prologue: weights (forward pass), initial iterate -------------------------------------
    if (tid == 0) { s_nan = 0; s_dead = 0; }
    __syncthreads();
    for (int i = beg + tid; i < end; i += nthr) {
        if (OP == 3) {
            for (int j = 0; j < K; ++j) {
                const int pred = a.idx[(size_t)j * cnt + i];
                double acc = 0.0;
                for (int aa = 0; aa < A; ++aa)
                    acc = fma(__ldg(a.p + ((size_t)aa * K + j) * cnt + i), a.policy_in[(size_t)pred * A + aa], acc);
                const double wj = a.term[pred] ? 0.0 : acc;
                a.w[(size_t)j * cnt + i] = wj;
                if (i - beg < nres) {
                    s_idx[j * nres + (i - beg)] = pred;
                    s_w[j * nres + (i - beg)] = wj;
                }
            }
            if (i - beg < nres) s_p0[i - beg] = a.c0[i];
        }
        put(0u, i, OP == kOpSoftVI ? kNegHuge : 0.0);
    }
    __syncthreads();
    if (tid == 0) publish(1ull);

    // one sweep: iterate number q (buffer q & 1) -> q + 1.  Votes of the CTA go to (any_gt, any_nan) on thread 0.
    auto sweep = [&](const unsigned q, const bool take_snap, bool &any_gt, bool &any_nan) -> bool {
        bool ok = true;
        const unsigned long long need = (unsigned long long)q + 1ull;
        for (int d = d_lo + tid; d <= d_hi; d += nthr)
            if (d != cta) ok = spin_ge(progress + (size_t)d * kProgressStride, need, false) && ok;
        if (__syncthreads_or(ok ? 0 : 1)) return false;
        const double *x_in = (q & 1u) ? buf1 : buf0;
        bool gt = false, nan = false;
        auto body = [&](auto xl) {
            if (OP == 3 && K_T > 0) {
                constexpr int KK = K_T > 0 ? K_T : 1, UU = U > 0 ? U : 1;
                for (int i0 = beg + tid; i0 < end; i0 += UU * nthr) {
                    int ix[UU][KK];
                    double wv[UU][KK], p0v[UU], xo[UU], xv[UU][KK];
#pragma unroll
                    for (int u = 0; u < UU; ++u) {
                        const int i = min(i0 + u * nthr, end - 1);      // clamped: inactive slots repeat a valid state
                        const int li = i - beg;
                        if (li < nres) {                                // resident rows (shared memory)
#pragma unroll
                            for (int j = 0; j < KK; ++j) {
                                ix[u][j] = s_idx[j * nres + li];
                                wv[u][j] = s_w[j * nres + li];
                            }
                            p0v[u] = s_p0[li];
                        } else {
#pragma unroll
                            for (int j = 0; j < KK; ++j) {
                                ix[u][j] = __ldg(a.idx + (size_t)j * cnt + i);
                                wv[u][j] = __ldg(a.w + (size_t)j * cnt + i);
                            }
                            p0v[u] = __ldg(a.c0 + i);
                        }
                        xo[u] = ld_cg(x_in + lo + i);
                    }
                    // Only the thread that OWNS a state may read its neighbours: a mailbox slot is overwritten as
                    // soon as the value computed from it has crossed back (see load_x), so a second reader -- a
                    // clamped, inactive slot -- could wait for a tag that is already gone.
#pragma unroll
                    for (int u = 0; u < UU; ++u)
#pragma unroll
                        for (int j = 0; j < KK; ++j) xv[u][j] = (i0 + u * nthr < end) ? xl(ix[u][j]) : 0.0;
#pragma unroll
                    for (int u = 0; u < UU; ++u) {
                        const int i = i0 + u * nthr;
                        if (i < end) {
                            double acc = 0.0;
#pragma unroll
                            for (int j = 0; j < KK; ++j) acc = fma(wv[u][j], xv[u][j], acc);   // same chain as slab_update
                            const double x = p0v[u] + acc;
                            const double diff = fabs(x - xo[u]);
                            gt |= diff > a.eps;
                            nan |= diff != diff;
                            if (take_snap) snap[i] = xo[u];
                            put(q + 1u, i, x);
                        }
                    }
                }
            } else {
                for (int i = beg + tid; i < end; i += nthr) {
                    const double x = CODED ? slab_update_p<OP, A_T, K_T>(a, [&](int aa, int j) {
                                                  return s_dict[__ldg(a.code + ((size_t)aa * K + j) * cnt + i)]; }, xl, i, nullptr)
                                           : slab_update<OP, A_T, K_T>(a, xl, i, nullptr);
                    const double xo = ld_cg(x_in + lo + i);
                    const double diff = fabs(x - xo);
                    gt |= diff > a.eps;
                    nan |= diff != diff;
                    if (take_snap) snap[i] = xo;
                    put(q + 1u, i, x);
                }
            }
        };
        if (near_lo || near_hi) body([&](int g) { return load_x(x_in, q, g); });
        else body([&](int g) { return ld_cg(x_in + g); });
        if (nan) s_nan = 1;
        const int any = __syncthreads_or(gt ? 1 : 0);
        if (s_dead) return false;
        if (tid == 0) {
            publish((unsigned long long)q + 2ull);
            any_gt = any != 0;
            any_nan = s_nan != 0;
            s_nan = 0;
        }
        return true;
    };

    // ---- chunks of sweeps ------------------------------------------------------------------------
    const int limit = a.max_sweeps > 0 ? a.max_sweeps : 0x7fffffff;
    int n = 0, st = IRLB200_ST_CONVERGED;
    unsigned q = 0;
    for (;;) {
        const int k = min(chunk, limit - n);
        unsigned long long my_gt = 0ull, my_nan = 0ull;
        bool ok = true;
        for (int i = 0; i < k && ok; ++i) {
            bool g = false, nn = false;
            ok = sweep(q, i == 0, g, nn);
            if (ok) {
                ++q;
                if (g) my_gt |= 1ull << i;
                if (nn) my_nan |= 1ull << i;
            }
        }
        if (!ok) dead = true;
        unsigned long long all_gt = 0ull, all_nan = 0ull;
        if (!full_barrier(my_gt, my_nan, all_gt, all_nan)) { st = IRLB200_ST_ABORTED; break; }
        // first sweep of the chunk that ends the reference's loop: delta <= eps, or NaN
        const unsigned long long kmask = k >= 64 ? ~0ull : ((1ull << k) - 1ull);
        const unsigned long long ends = (~all_gt | all_nan) & kmask;
        if (!ends) {
            n += k;
            if (n >= limit) { st = IRLB200_ST_MAXSWEEPS; break; }
            continue;
        }
        const int stop = __ffsll((long long)ends) - 1;
        st = ((all_nan >> stop) & 1ull) ? IRLB200_ST_NONFINITE : IRLB200_ST_CONVERGED;
        if (stop == k - 1) { n += k; break; }                 // ended on the chunk's last sweep: nothing to undo
        // Restore the iterate the chunk started from and replay sweeps 0..stop.  The iterate number jumps
        // by 2 (same buffer parity) so that the mailbox tags of the restored rows differ from the tags
        // of the rows they replace; everybody is behind the barrier above, the progress words order the rest.
        q += 2u;
        for (int i = beg + tid; i < end; i += nthr) put(q, i, snap[i]);
        __syncthreads();
        if (tid == 0) publish((unsigned long long)q + 1ull);
        for (int i = 0; i <= stop && ok; ++i) {
            bool g, nn;
            ok = sweep(q, false, g, nn);
            if (ok) ++q;
        }
        if (!ok) { st = IRLB200_ST_ABORTED; break; }
        n += stop + 1;
        break;
    }

    // ---- outputs ---------------------------------------------------------------------------------
    if (st != IRLB200_ST_ABORTED) {
        const double *x_new = (q & 1u) ? buf1 : buf0, *x_old = ((q - 1u) & 1u) ? buf1 : buf0;
        for (int i = beg + tid; i < end; i += nthr) {
            a.out[i] = ld_cg(x_new + lo + i);
            if (OP == kOpSoftVI && a.policy_out && q > 0) {
                double qv[QN];
                auto xl = [&](int g) { return load_x(x_old, q - 1u, g); };
                const double x = CODED ? slab_update_p<OP, A_T, K_T>(a, [&](int aa, int j) {
                                              return s_dict[__ldg(a.code + ((size_t)aa * K + j) * cnt + i)]; }, xl, i, qv)
                                       : slab_update<OP, A_T, K_T>(a, xl, i, qv);
                for (int aa = 0; aa < A; ++aa) a.policy_out[(size_t)i * A + aa] = exp(qv[aa] - x);     // maxent.py:341
            }
        }
    }
    if (cta == 0 && tid == 0) {
        if (n_iter) *n_iter = n;
        if (status) *status = st;
    }
}

}  // namespace irlb200

using namespace irlb200;

static int flow_env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// bytes of the local (not peer-mapped) work buffer of irlb200_slab_flow: progress words + chunk snapshot
extern "C" size_t irlb200_slab_flow_work_bytes(int cnt) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t prog = (size_t)sms * 8 * kProgressStride * sizeof(unsigned long long);
    return prog + sizeof(double) * (size_t)(cnt > 0 ? cnt : 0) + 256;
}

// peer-mapped block of a rank for irlb200_slab_flow (also valid for irlb200_slab_persistent):
// [header | iterate 0 | iterate 1 | LL mailboxes for one ghost row from each neighbour, double-buffered]
extern "C" size_t irlb200_slab_flow_block_bytes(int S_total, int halo) {
    return flow_mail_offset(S_total) + 4 * sizeof(LLSlot) * (size_t)(halo > 0 ? halo : 0);
}

// zero the header and the mailboxes (all ranks, before any rank launches)
extern "C" int irlb200_slab_flow_reset(void *block, int S_total, int halo, void *stream) {
    if (!block || S_total <= 0 || halo < 0) return fail(IRLB200_EINVAL, "slab_flow_reset: bad argument");
    cudaError_t e = cudaMemsetAsync(block, 0, kSlabHeaderBytes, (cudaStream_t)stream);
    if (e == cudaSuccess && halo > 0)
        e = cudaMemsetAsync(static_cast<unsigned char *>(block) + flow_mail_offset(S_total), 0,
                            4 * sizeof(LLSlot) * (size_t)halo, (cudaStream_t)stream);
    return e == cudaSuccess ? IRLB200_OK : fail_cuda(e, "cudaMemsetAsync(slab flow header)");
}

extern "C" int irlb200_slab_flow(int op, int rank, int world, void *const *blocks, int S_total, int lo, int cnt,
                                 int halo, int A, int K, const int32_t *idx, const double *p, const double *c0,
                                 const double *c1, const double *policy_in, const uint8_t *terminal_mask,
                                 double *w_scratch, double discount, double eps, int max_sweeps, int vi_mean,
                                 double *out, double *policy_out, int32_t *n_iter, int32_t *status, double timeout_s,
                                 int chunk, void *work, size_t work_bytes, void *stream) {
    return irlb200_slab_flow_coded(op, rank, world, blocks, S_total, lo, cnt, halo, A, K, idx, p, nullptr, nullptr, c0, c1,
                                   policy_in, terminal_mask, w_scratch, discount, eps, max_sweeps, vi_mean, out,
                                   policy_out, n_iter, status, timeout_s, chunk, work, work_bytes, stream);
}

extern "C" int irlb200_slab_flow_coded(int op, int rank, int world, void *const *blocks, int S_total, int lo, int cnt,
                                       int halo, int A, int K, const int32_t *idx, const double *p,
                                       const uint8_t *p_code, const double *p_dict, const double *c0,
                                       const double *c1, const double *policy_in, const uint8_t *terminal_mask,
                                       double *w_scratch, double discount, double eps, int max_sweeps, int vi_mean,
                                       double *out, double *policy_out, int32_t *n_iter, int32_t *status,
                                       double timeout_s, int chunk, void *work, size_t work_bytes, void *stream) {
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !blocks || cnt <= 0 || halo <= 0 || !idx || !p ||
        !c0 || !out || !work)
        return fail(IRLB200_EINVAL, "slab_flow: bad argument");
    if (op < 1 || op > 3) return fail(IRLB200_EINVAL, "slab_flow: unknown op");
    if (device_count_impl() <= 0) return fail(IRLB200_ECUDA, "no CUDA device");
    if (!(A == 4 && (K == 5 || K == 4)) && A > kMaxDynA) return fail(IRLB200_EINVAL, "run-time A > 16 is not supported");
    if (op == 3 && (!policy_in || !terminal_mask || !w_scratch)) return fail(IRLB200_EINVAL, "slab_flow: forward pass inputs");
    if (op == 1 && (!c1 || !policy_out)) return fail(IRLB200_EINVAL, "slab_flow: soft-VI inputs");
    if (world > 1 && cnt < halo) return fail(IRLB200_ELIMIT, "slab_flow: a slab must hold at least one grid row");
    if (work_bytes < irlb200_slab_flow_work_bytes(cnt)) return fail(IRLB200_EINVAL, "slab_flow: work buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    SlabPeers pe{};
    for (int r = 0; r < world; ++r) pe.shared[r] = reinterpret_cast<SlabShared *>(blocks[r]);
    pe.rank = rank; pe.world = world; pe.lo = lo; pe.hi = lo + cnt; pe.halo = halo;
    pe.timeout_ns = (long long)((timeout_s > 0 ? timeout_s : 20.0) * 1e9);
    unsigned char *base = static_cast<unsigned char *>(blocks[rank]);

    OverlapArgs oa{};
    oa.op = op; oa.lo = lo; oa.cnt = cnt; oa.S_total = S_total; oa.halo = halo; oa.A = A; oa.K = K;
    oa.idx = idx; oa.p = p; oa.c0 = c0; oa.c1 = c1; oa.policy_in = policy_in; oa.term = terminal_mask;
    oa.w = w_scratch; oa.discount = discount; oa.eps = eps; oa.max_sweeps = max_sweeps; oa.vi_mean = vi_mean;
    oa.out = out; oa.policy_out = policy_out;
    const bool coded = p_code && p_dict && op != 3 && A == 4 && (K == 4 || K == 5) && flow_env_int("IRLB200_FLOW_CODED", 1);
    oa.code = coded ? p_code : nullptr; oa.dict = coded ? p_dict : nullptr;

    const bool fast = (A == 4 && K == 5), compact = (A == 4 && K == 4);
    const void *k = nullptr;
    // forward pass: states per thread in flight (U) x CTAs per SM; IRLB200_FLOW_FWD = 14 | 24 | 42 | 23 (U, CTAs/SM)
    const int fwd = flow_env_int("IRLB200_FLOW_FWD", 24);
    if (op == 3) {
        if (fast) k = fwd == 14 ? (const void *)slab_flow_kernel<3, 4, 5, 1, 4> : fwd == 42 ? (const void *)slab_flow_kernel<3, 4, 5, 4, 2>
                    : fwd == 23 ? (const void *)slab_flow_kernel<3, 4, 5, 2, 3> : (const void *)slab_flow_kernel<3, 4, 5, 2, 4>;
        else if (compact) k = fwd == 14 ? (const void *)slab_flow_kernel<3, 4, 4, 1, 4> : fwd == 42 ? (const void *)slab_flow_kernel<3, 4, 4, 4, 2>
                    : fwd == 23 ? (const void *)slab_flow_kernel<3, 4, 4, 2, 3> : (const void *)slab_flow_kernel<3, 4, 4, 2, 4>;
        else k = (const void *)slab_flow_kernel<3, 0, 0, 1, 4>;
    } else if (op == 1) {
        // coded rows: the sweep is FP64 bound (exp / log), and four CTAs of 256 threads per SM (64 registers, a small
        // spill) hide its latencies better than three: 53.2 -> 47.0 us per sweep at 2.1 M states (2: 59.7)
        const int occ = flow_env_int("IRLB200_FLOW_SOFTVI_OCC", 4);
        if (coded && occ == 4) k = fast ? (const void *)slab_flow_kernel<kOpSoftVI, 4, 5, 1, 4, true> : (const void *)slab_flow_kernel<kOpSoftVI, 4, 4, 1, 4, true>;
        else if (coded) k = fast ? (const void *)slab_flow_kernel<kOpSoftVI, 4, 5, 1, 3, true> : (const void *)slab_flow_kernel<kOpSoftVI, 4, 4, 1, 3, true>;
        else k = fast ? (const void *)slab_flow_kernel<kOpSoftVI, 4, 5, 1, 3> : compact ? (const void *)slab_flow_kernel<kOpSoftVI, 4, 4, 1, 3>
                      : (const void *)slab_flow_kernel<kOpSoftVI, 0, 0, 1, 3>;
    } else {
        if (coded) k = fast ? (const void *)slab_flow_kernel<kOpVI, 4, 5, 1, 3, true> : (const void *)slab_flow_kernel<kOpVI, 4, 4, 1, 3, true>;
        else k = fast ? (const void *)slab_flow_kernel<kOpVI, 4, 5, 1, 3> : compact ? (const void *)slab_flow_kernel<kOpVI, 4, 4, 1, 3>
                      : (const void *)slab_flow_kernel<kOpVI, 0, 0, 1, 3>;
    }

    const int threads = 256;
    int dev = 0, sms = 0, per_sm = 0, nb = 0, edge = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // sweeps per stop-rule chunk: the forward pass runs 10^4-10^7 sweeps, so the longest chunk wins (4.50 -> 4.37 us per
    // sweep on a 524 k-state slab); soft-VI / VI converge within ~10^3 sweeps, where a chunk of 64 wastes up to 3 % in the
    // final undo + replay
    if (chunk <= 0) chunk = flow_env_int("IRLB200_FLOW_CHUNK", op == 3 ? 64 : 32);
    if (chunk > 64) chunk = 64;
    cudaError_t e = cudaSuccess;
    // grid of a kernel variant: all CTAs co-resident (cooperative launch), one CTA per >= 256 states
    auto plan = [&](const void *kern) -> int {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
        if (e != cudaSuccess) return fail_cuda(e, "occupancy(slab flow)");
        if (per_sm < 1) return fail(IRLB200_ELIMIT, "slab flow kernel does not fit on an SM");
        if (per_sm > 8) per_sm = 8;
        const int per_sm_env = flow_env_int("IRLB200_FLOW_CTAS_PER_SM", 0);
        if (per_sm_env > 0 && per_sm_env < per_sm) per_sm = per_sm_env;
        long long want = ((long long)cnt + threads - 1) / threads, cap = (long long)sms * per_sm;
        nb = (int)(want < cap ? want : cap);
        const int nb_env = flow_env_int("IRLB200_FLOW_BLOCKS", 0);
        if (nb_env > 0 && nb_env < nb) nb = nb_env;
        if (nb < 1) nb = 1;
        // CTAs per boundary row (see the kernel): about one state per thread, at most 8
        edge = halo / threads;
        edge = edge < 1 ? 1 : (edge > 8 ? 8 : edge);
        edge = flow_env_int("IRLB200_FLOW_EDGE_CTAS", edge);
        if (world == 1 || edge < 0 || cnt < 4 * halo || nb < 8 * edge || halo < edge) edge = 0;
        return IRLB200_OK;
    };

    // Forward pass with compile-time K: if the table rows of EVERY CTA's range fit its share of the SM's shared
    // memory, keep them there for the whole fixed point (2048 x 2048 on 8 GPUs: 903 states x 56 B per CTA,
    // 4 CTAs per SM; measured 4.77 -> 4.23 us per sweep on that slab).  Partial residency was measured and
    // dropped: the large carve-out shrinks L1 and the streamed rest gets slower (23.9 -> 28.5 us at 2 M states).
    size_t dyn = 0;
    int res_cap = 0;
    if (op == 3 && (fast || compact) && flow_env_int("IRLB200_FLOW_RESIDENT", 1) && fwd == 24) {
        const void *k1 = fast ? (const void *)slab_flow_kernel<3, 4, 5, 1, 4> : (const void *)slab_flow_kernel<3, 4, 4, 1, 4>;
        if (int rc = plan(k1)) return rc;
        const int n_edge = (rank > 0 ? edge : 0) + (rank < world - 1 ? edge : 0);
        const long long len_max = edge ? ((long long)cnt - (long long)halo * ((rank > 0) + (rank < world - 1))) / (nb - n_edge) + 2
                                       : ((long long)cnt + nb - 1) / nb + 1;
        int smem_sm = 0, smem_optin = 0, reserved = 1024;
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev);
        long long share = (long long)smem_sm / per_sm - reserved - 256;           // 256: the kernel's static shared memory
        if (share > smem_optin) share = smem_optin;
        share &= ~127ll;
        const long long need = len_max * (12 * K + 8) + 32;
        const long long rows_halo = edge ? ((long long)halo + edge - 1) / edge + 1 : 0;   // an edge CTA's range
        if (share >= need && share >= rows_halo * (12 * K + 8) + 32) {
            int fit = 0;
            e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)share);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, k1, threads, (size_t)share);
            if (e == cudaSuccess && fit >= per_sm) {
                k = k1;
                dyn = (size_t)share;
                res_cap = (int)((dyn - 16) / (size_t)(12 * K + 8));
            }
            cudaGetLastError();
        }
    }
    if (!dyn)
        if (int rc = plan(k)) return rc;

    unsigned long long *progress = static_cast<unsigned long long *>(work);
    const size_t prog_bytes = (size_t)sms * 8 * kProgressStride * sizeof(unsigned long long);
    double *snap = reinterpret_cast<double *>(static_cast<unsigned char *>(work) + ((prog_bytes + 255) & ~(size_t)255));
    e = cudaMemsetAsync(progress, 0, prog_bytes, st);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync(slab flow progress)");
    void *params[10] = {&oa, &pe, &base, &progress, &snap, &chunk, &edge, &res_cap, &n_iter, &status};
    e = cudaLaunchCooperativeKernel(k, dim3(nb), dim3(threads), params, dyn, st);
    if (e != cudaSuccess) return fail_cuda(e, "cudaLaunchCooperativeKernel(slab flow)");
    return IRLB200_OK;
}
