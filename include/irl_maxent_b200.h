/*
 * irl_maxent_b200.h -- C ABI of libirlmaxent_b200.so
 *
 * B200-native (sm_100a) replacement for the arithmetic behind the MaxEnt /
 * MaxCausalEnt IRL hot path of narendasan/irl-maxent.  The reference has no FFI
 * layer -- its boundary is a set of Python functions on dense numpy arrays
 * (src/maxent.py, src/solver.py).  Each entry point below names the reference
 * function (file:line under /root/reference) whose arithmetic it replaces; the
 * ctypes stub a maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless its name starts with `h_`.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  All
 *     work is enqueued on it; nothing synchronises the host unless documented.
 *     Internal scratch (barrier words and iterate buffers of the cooperative-grid
 *     mode, weight scratch of the streamed forward pass) is cached per
 *     (device, stream): calls on different streams never share it, calls on one
 *     stream are ordered by the stream.  Entry points that take a caller-owned
 *     work buffer (irlb200_slab_flow) use no internal scratch at all.
 *   - Every function returns 0 on success, a negative IRLB200_E* code otherwise;
 *     irlb200_last_error() gives a message for the calling thread.
 *   - Values are IEEE float64 (the reference computes in float64 only);
 *     indices are int32 (S < 2^31).
 *   - There is NO CPU fallback: without a CUDA device every compute entry
 *     point returns IRLB200_ECUDA.
 *
 * Table layout ("state-merged ELL", slot-major so that threads mapped to
 * consecutive states read consecutive addresses):
 *   successors   succ_idx[K][S]  int32    j-th distinct successor of state s
 *                succ_p  [A][K][S] f64    P[s, succ_idx[j][s], a]
 *   predecessors pred_idx[K][S]  int32    j-th distinct predecessor of state s'
 *                pred_p  [A][K][S] f64    P[pred_idx[j][s'], s', a]
 *   Slots hold the distinct neighbours in ascending state order; unused slots
 *   are padded with idx = own state, p = 0.  K is discovered from the data
 *   (irlb200_dense_count), not assumed.
 *   A batch of B problems stores B such tables back to back (`table_stride`
 *   elements apart, counted in states*slots, see each call); stride 0 shares
 *   one table between all problems of the batch.
 */
#ifndef IRL_MAXENT_B200_H
#define IRL_MAXENT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRLB200_OK        0
#define IRLB200_EINVAL   (-1)   /* bad argument (shape, null pointer, unsupported size) */
#define IRLB200_ECUDA    (-2)   /* CUDA runtime error / no device */
#define IRLB200_ELIMIT   (-3)   /* problem does not fit the requested execution mode */

/* status bits written to `status` outputs by the sweep kernels */
#define IRLB200_ST_CONVERGED   0   /* stopped because delta <= eps (or fixed count reached) */
#define IRLB200_ST_NONFINITE   1   /* stopped because delta was NaN (reference: NaN ends the loop) */
#define IRLB200_ST_MAXSWEEPS   2   /* stopped by the max_sweeps guard (reference would keep looping) */
#define IRLB200_ST_ABORTED     3   /* slab mode: a peer rank did not show up within the timeout */

/* execution modes for a single problem (batched calls always use IRLB200_MODE_CTA) */
#define IRLB200_MODE_AUTO     0
#define IRLB200_MODE_CTA      1   /* one CTA owns the problem: iterate in shared memory      */
#define IRLB200_MODE_CLUSTER  2   /* one thread-block cluster (<= 16 CTAs) owns it: iterate in registers +
                                     distributed shared memory, rows and votes exchanged by st.async into the
                                     peers' shared memory.  Exists for irlb200_backward and irlb200_svf on 5-slot
                                     grid-stencil tables with n <= 128, n % 4 == 0 (also batches of such worlds);
                                     IRLB200_ELIMIT otherwise.  AUTO picks it for worlds beyond one CTA.        */
#define IRLB200_MODE_GRID     3   /* cooperative persistent grid: iterate in L2/HBM          */

int         irlb200_version(void);
const char *irlb200_last_error(void);
/* number of CUDA devices visible to the library (0 => every compute call fails) */
int         irlb200_device_count(void);
/* limits of the execution modes on the current device */
int         irlb200_max_states_cta(void);
int         irlb200_max_states_cluster(void);

/* ------------------------------------------------------------------------- *
 * (1) one-time compression of the dense table
 *     replaces the per-call dense slicing `[np.array(p_transition[:, :, a]) ...]`
 *     maxent.py:102,143,320 and solver.py:37, and the copy+mask maxent.py:98-99.
 * ------------------------------------------------------------------------- */

/* Pass 1: count distinct successors / predecessors of every state.
 *   P        [S][S][A] f64, C-contiguous (A innermost: gridworld.py:124-142)
 *   succ_cnt [S] int32 out, pred_cnt [S] int32 out
 *   kmax     [2] int32 out: {max succ_cnt, max pred_cnt}                     */
int irlb200_dense_count(const double *P, int S, int A,
                        int32_t *succ_cnt, int32_t *pred_cnt, int32_t *kmax, void *stream);

/* Pass 2: fill the tables (layout above) for slot counts Ks >= kmax[0], Kp >= kmax[1].
 *   pred_cnt is consumed as scratch (must hold the pass-1 values, returns them unchanged). */
int irlb200_dense_fill(const double *P, int S, int A, int Ks, int Kp,
                       int32_t *succ_idx, double *succ_p,
                       int32_t *pred_idx, double *pred_p,
                       int32_t *pred_cnt, void *stream);

/* Tables of GridWorld / IcyGridWorld built directly, no dense detour
 *   (gridworld.py:144-171 deterministic when icy == 0, :200-248 icy otherwise).
 *   Ks = Kp = 5.  B worlds with per-world slip probability p_slip[b] (device, f64);
 *   tables of world b start at element b * (5*S) of the idx arrays and
 *   b * (A*5*S) of the p arrays.  Values are bit-identical to the reference's. */
int irlb200_gridworld_tables(int size, int icy, int B, const double *p_slip,
                             int32_t *succ_idx, double *succ_p,
                             int32_t *pred_idx, double *pred_p, void *stream);

/* Rows of the states [lo, lo+cnt) of ONE world only: arrays of stride cnt holding GLOBAL neighbour
 * indices (slab mode: every rank builds just its own rows; 2048x2048 never exists densely). */
int irlb200_gridworld_tables_range(int size, int icy, double p_slip, int lo, int cnt,
                                   int32_t *succ_idx, double *succ_p,
                                   int32_t *pred_idx, double *pred_p, void *stream);

/* The same builders with K slots per state instead of 5.  K = 4 is the compact form: a grid-world
 * state has at most 4 distinct successors / predecessors (4 neighbours, or 3 + itself on an edge),
 * so no entry is lost, the streamed (HBM-bound) kernels move 15-22 % fewer bytes per sweep, and
 * results are bitwise those of the 5-slot tables (a padding slot only ever adds fma(0, x, acc)).
 * The register-resident / tiled / cluster kernels need the 5-slot form. */
int irlb200_gridworld_tables_k(int size, int icy, int B, const double *p_slip, int K,
                               int32_t *succ_idx, double *succ_p,
                               int32_t *pred_idx, double *pred_p, void *stream);
int irlb200_gridworld_tables_range_k(int size, int icy, double p_slip, int lo, int cnt, int K,
                                     int32_t *succ_idx, double *succ_p,
                                     int32_t *pred_idx, double *pred_p, void *stream);

/* Dense P[S][S][A] of the same worlds, written on the device (test helper for the
 * compression kernels at sizes where the Python table builder is too slow). */
int irlb200_gridworld_dense(int size, int icy, double p_slip, double *P, void *stream);

/* ------------------------------------------------------------------------- *
 * Problem descriptor shared by the sweep entry points
 * ------------------------------------------------------------------------- */
typedef struct irlb200_tables {
    int32_t S, A;             /* states, actions                                         */
    int32_t Ks, Kp;           /* slots per state in the successor / predecessor table    */
    const int32_t *succ_idx;  /* [B or 1][Ks][S]                                          */
    const double  *succ_p;    /* [B or 1][A][Ks][S]                                       */
    const int32_t *pred_idx;  /* [B or 1][Kp][S]                                          */
    const double  *pred_p;    /* [B or 1][A][Kp][S]                                       */
    int32_t shared;           /* 1: one table for the whole batch, 0: one table per problem */
    int32_t stencil_n;        /* > 0: the caller asserts that S = n*n and every table entry with
                                 non-zero probability links s to one of s-n, s-1, s, s+1, s+n without
                                 wrapping around a row end (any GridWorld / IcyGridWorld; checked by
                                 the host layer when tables are built).  Enables the stencil-tiled
                                 kernels; 0 = no assumption, generic gather.                        */
} irlb200_tables;

/* ------------------------------------------------------------------------- *
 * (2) non-causal backward pass -- local_action_probabilities, maxent.py:119-159
 *     fused: exp(reward) (:142), n_sweeps partition sweeps (:154-156, the
 *     reference uses 2*S), action normalisation (:159).  Range-extended by an
 *     exact power-of-two rescale (bit-identical to the raw loop where that is
 *     finite).
 *   reward   [B][S]          terminal_mask [B or 1][S] uint8 (1 = terminal; zs seed :146-147)
 *   policy   [B][S][A] out
 * ------------------------------------------------------------------------- */
int irlb200_backward(const irlb200_tables *t, int B, const double *reward,
                     const uint8_t *terminal_mask, int mask_shared, int n_sweeps,
                     double *policy, int mode, void *stream);

/* ------------------------------------------------------------------------- *
 * (3) causal soft value iteration -- local_causal_action_probabilities,
 *     maxent.py:279-341 (softmax fold :260-276, -1e200 start :323, policy :341)
 *   phi [B or 1][S] terminal reward function (0 at terminals, -inf elsewhere, or
 *       the caller's array, :312-317)
 *   policy [B][S][A] out, n_iter [B] int32 out, status [B] int32 out
 *   max_sweeps <= 0 means no guard (the reference has none).
 * ------------------------------------------------------------------------- */
int irlb200_soft_vi(const irlb200_tables *t, int B, const double *reward,
                    const double *phi, int phi_shared, double discount, double eps,
                    int max_sweeps, double *policy, double *value_out /* [B][S] or NULL */,
                    int32_t *n_iter, int32_t *status, int mode, void *stream);

/* ------------------------------------------------------------------------- *
 * value iteration -- solver.value_iteration, solver.py:9-52
 *   kind 0: max over actions (:47); kind 1: mean over actions
 *   (stochastic_value_iteration, solver.py:55-104)
 * ------------------------------------------------------------------------- */
int irlb200_value_iteration(const irlb200_tables *t, int B, const double *reward,
                            double discount, double eps, int max_sweeps, int kind,
                            double *value, int32_t *n_iter, int32_t *status,
                            int mode, void *stream);

/* ------------------------------------------------------------------------- *
 * (4) forward state-visitation pass -- expected_svf_from_policy, maxent.py:63-114
 *     gather over predecessors; the outgoing rows of terminal states are
 *     dropped (:98-99) through the mask; per-policy predecessor weights
 *     W[j][s'] = sum_a P[s_j,s',a] * policy[s_j,a] are formed on chip.
 *   p_initial [B or 1][S], policy [B][S][A], svf [B][S] out
 *   grad      [B][S] out or NULL: fused epilogue for identity features,
 *             grad = e_features - svf   (maxent.py:248 with features = I);
 *             e_features [B or 1][S] (ef_shared = 1: one vector for the whole batch)
 * ------------------------------------------------------------------------- */
int irlb200_svf(const irlb200_tables *t, int B, const double *p_initial, int p0_shared,
                const uint8_t *terminal_mask, int mask_shared, const double *policy,
                double eps, int max_sweeps,
                double *svf, const double *e_features, int ef_shared, double *grad,
                int32_t *n_iter, int32_t *status, int mode, void *stream);

/* The same call with a launch-order hint for batches: `order` [B] int32 (device) is a permutation of
 * 0..B-1, block i of the one-CTA-per-problem launch works on problem order[i] (NULL: identity).
 * Results are independent of the order; sorting the problems longest-first (e.g. by the sweep
 * counts `n_iter` of the previous gradient step) removes the tail of a batch whose forward passes
 * differ in length by several x.  Ignored by the cluster / cooperative-grid modes. */
int irlb200_svf_ordered(const irlb200_tables *t, int B, const double *p_initial, int p0_shared,
                        const uint8_t *terminal_mask, int mask_shared, const double *policy,
                        double eps, int max_sweeps,
                        double *svf, const double *e_features, int ef_shared, double *grad,
                        int32_t *n_iter, int32_t *status, int mode, const int32_t *order, void *stream);

/* ------------------------------------------------------------------------- *
 * fused gradient step: (2)+(4) or (3)+(4) in ONE launch per batch --
 * compute_expected_svf maxent.py:162-193 / compute_expected_causal_svf :344-380,
 * the policy never leaves the SM.  causal == 0: backward pass with n_backward
 * sweeps; causal != 0: soft-VI with (discount, eps_lap).
 *   n_iter [B][2] int32 out: {policy sweeps, svf sweeps};  status [B][2]
 * ------------------------------------------------------------------------- */
int irlb200_expected_svf(const irlb200_tables *t, int B, int causal,
                         const double *reward, const double *p_initial, int p0_shared,
                         const uint8_t *terminal_mask, const double *phi, int mask_shared,
                         int n_backward, double discount, double eps_lap, double eps_svf,
                         int max_sweeps, double *svf, const double *e_features, int ef_shared,
                         double *grad, double *policy_out /* [B][S][A] or NULL */,
                         int32_t *n_iter, int32_t *status, void *stream);

/* ------------------------------------------------------------------------- *
 * the whole outer gradient loop on the device -- `while delta > eps` of irl (maxent.py:240-252) and
 * irl_causal (:436-450) -- for tiny worlds (S <= 32, A = 4, 5-slot tables: BASELINE configs[0..1]) with
 * identity features and a built-in optimizer: one warp per problem runs, per step, the gradient-step body
 * on reward = omega (:244), grad = e_features - svf (:248), the update rule and the delta test (:252).
 *   theta  [B][S] in / out (omega);   e_features [B or 1][S]
 *   opt_kind 0: Sga, omega += lr * grad (optimizer.py:104-107); 1: ExpSga, omega *= exp(lr * grad) (:161-164)
 *   lr     [B or 1][n_rates] (device): the learning rates of the next n_rates steps -- the schedule
 *          (optimizer.py:217-293) stays a host-side callable, the host evaluates it ahead
 *   steps  [B] in / out, accumulated;  done [B] in / out: 1 = delta <= eps or NaN reached (problems that are
 *          done are skipped); a problem that is not done after n_rates steps needs another call with the
 *          next rates.  last_counts [B][2] or NULL: {policy sweeps, forward sweeps} of the last step.
 * ------------------------------------------------------------------------- */
int irlb200_irl_small(const irlb200_tables *t, int B, int causal, double *theta,
                      const double *e_features, int ef_shared, const double *p_initial, int p0_shared,
                      const uint8_t *terminal_mask, const double *phi, int mask_shared,
                      int n_backward, double discount, double eps_lap, double eps_svf, int max_sweeps,
                      int opt_kind, const double *lr, int lr_shared, int n_rates, double eps,
                      int32_t *steps, int32_t *done, int32_t *last_counts, void *stream);

/* ------------------------------------------------------------------------- *
 * slab mode (one huge MDP sharded by contiguous state ranges over several GPUs; the halo
 * exchange and the convergence all-reduce run between launches, see irl-maxent_b200/slab.py).
 * ONE sweep over the owned states [lo, lo+cnt) of a full-length iterate:
 *   op 1: soft-VI sweep (maxent.py:329-338), p = succ_p [A][K][cnt], c0 = reward, c1 = phi
 *   op 2: value-iteration sweep (solver.py:44-50),  p = succ_p, c0 = reward
 *   op 3: forward sweep (maxent.py:109-112),        p = W [K][cnt] from irlb200_slab_weights, c0 = p_initial
 *   idx [K][cnt] holds global indices; c0/c1/policy are local (length cnt); x_in / x_out are
 *   full-length, only x_out[lo..lo+cnt) is written.  vote[0] |= (some |diff| > eps),
 *   vote[1] |= (some diff is NaN).  policy ([cnt][A] or NULL, op 1): exp(q - v) of this sweep (:341).
 * ------------------------------------------------------------------------- */
int irlb200_slab_sweep(int op, int lo, int cnt, int A, int K, const int32_t *idx, const double *p,
                       const double *c0, const double *c1, double discount, double eps, int vi_mean,
                       const double *x_in, double *x_out, int32_t *vote, double *policy, void *stream);
/* W[j][i] = sum_a pred_p[a][j][i] * policy[pred][a] (0 if pred terminal); policy [S_total][A] and
 * terminal_mask [S_total] are globally indexed (ghost rows exchanged by the caller). */
int irlb200_slab_weights(int cnt, int A, int K, const int32_t *pred_idx, const double *pred_p,
                         const double *policy, const uint8_t *terminal_mask, double *W, void *stream);

/* ------------------------------------------------------------------------- *
 * slab mode with the halo exchange inside the kernel (NVLink peer stores, no collective call).
 * Every rank owns a block of irlb200_slab_block_bytes(S_total) bytes from irlb200_peer_alloc
 * ([4 KiB barrier/flag header | iterate buffer 0 [S_total] | iterate buffer 1 [S_total]]),
 * exports it with irlb200_ipc_export (64-byte CUDA IPC handle, exchanged by the host layer over
 * torch.distributed) and maps the others with irlb200_ipc_import.  The header must be zero on all
 * ranks before any rank launches.  irlb200_slab_persistent runs one whole fixed point of the
 * slab [lo, lo+cnt) in ONE cooperative launch per rank:
 *   op 1 soft-VI  (idx/p successor rows [K][cnt] / [A][K][cnt], c0 reward, c1 phi; policy_out [cnt][A],
 *                  out = value [cnt])
 *   op 2 value iteration (c0 reward; out = value [cnt])
 *   op 3 forward pass (idx/p predecessor rows, c0 p_initial; policy_in [S_total][A] and
 *                  terminal_mask [S_total] globally indexed with valid ghost rows; w_scratch [K][cnt];
 *                  out = svf [cnt])
 * halo = ghost width in states (one grid row).  blocks[r] = base of rank r's block (world <= 16).
 * overlap: boundary-first kernel -- a group of CTAs per boundary row computes and pushes that row
 * first, so the system-scope fence and the NVLink flight overlap the interior sweep (same results).
 * 0 = never, 1 = for the ops where it pays (soft-VI, VI), 2 = always.
 * ------------------------------------------------------------------------- */
int    irlb200_peer_alloc(size_t bytes, void **ptr);
int    irlb200_peer_free(void *ptr);
int    irlb200_ipc_export(void *ptr, unsigned char *handle64);
int    irlb200_ipc_import(const unsigned char *handle64, void **ptr);
int    irlb200_ipc_close(void *ptr);
size_t irlb200_slab_block_bytes(int S_total);
int    irlb200_slab_reset(void *block, void *stream);      /* zero the header (all ranks, before a launch) */
int    irlb200_slab_persistent(int op, int rank, int world, void *const *blocks, int S_total, int lo,
                               int cnt, int halo, int A, int K, const int32_t *idx, const double *p,
                               const double *c0, const double *c1, const double *policy_in,
                               const uint8_t *terminal_mask, double *w_scratch, double discount,
                               double eps, int max_sweeps, int vi_mean, double *out, double *policy_out,
                               int32_t *n_iter, int32_t *status, double timeout_s, int overlap,
                               void *stream);

/* Slab mode without a per-sweep barrier (csrc/slab_flow.cu): same arguments and results as
 * irlb200_slab_persistent (blocks of irlb200_slab_flow_block_bytes bytes), but every persistent CTA
 * owns a fixed range of states and waits only for the CTAs (and, on the slab's first / last grid
 * row, the neighbouring GPU's mailbox) within one grid row of it; the stop rule is all-reduced once
 * per `chunk` sweeps (<= 64; <= 0: default, 64 for the forward pass, 32 otherwise) and the exact
 * stopping sweep of the reference (maxent.py:108,326; solver.py:40) is reproduced by
 * snapshot-and-replay inside the kernel.  Requirement (true for every GridWorld / IcyGridWorld slab): the
 * coupling across a slab boundary is one-to-one and symmetric -- a state s in the first / last `halo` states
 * of a slab links to s -+ halo in the neighbouring slab and to nothing else there, and that state links
 * back to s; other tables must use irlb200_slab_persistent.  `work` is a caller-owned device buffer of at least
 * irlb200_slab_flow_work_bytes(cnt) bytes, private to this call (not peer-mapped; any contents). */
size_t irlb200_slab_flow_work_bytes(int cnt);
/* The peer-mapped block of a rank for irlb200_slab_flow is the block of irlb200_slab_persistent followed by
 * "LL" mailboxes for one ghost row (halo states) from each neighbour: boundary-row values cross NVLink as
 * two 8-byte words carrying half the value and the iterate number each, so data and arrival are one one-way
 * store (no system-scope fence, no flag).  irlb200_slab_flow_reset zeroes header and mailboxes (all ranks,
 * before any rank launches); the block also serves irlb200_slab_persistent. */
size_t irlb200_slab_flow_block_bytes(int S_total, int halo);
int    irlb200_slab_flow_reset(void *block, int S_total, int halo, void *stream);
int    irlb200_slab_flow(int op, int rank, int world, void *const *blocks, int S_total, int lo, int cnt,
                         int halo, int A, int K, const int32_t *idx, const double *p, const double *c0,
                         const double *c1, const double *policy_in, const uint8_t *terminal_mask,
                         double *w_scratch, double discount, double eps, int max_sweeps, int vi_mean,
                         double *out, double *policy_out, int32_t *n_iter, int32_t *status,
                         double timeout_s, int chunk, void *work, size_t work_bytes, void *stream);

/* The same call with dictionary-coded successor probabilities for op 1 / 2: p_code [A][K][cnt] uint8 and p_dict [256]
 * f64 with p[a][j][i] == p_dict[p_code[a][j][i]] exactly (a grid world's table holds a handful of distinct values);
 * the sweeps then read one byte per table entry instead of eight -- bitwise the same results.  NULL codes: plain p. */
int    irlb200_slab_flow_coded(int op, int rank, int world, void *const *blocks, int S_total, int lo, int cnt,
                               int halo, int A, int K, const int32_t *idx, const double *p,
                               const uint8_t *p_code, const double *p_dict, const double *c0,
                               const double *c1, const double *policy_in, const uint8_t *terminal_mask,
                               double *w_scratch, double discount, double eps, int max_sweeps, int vi_mean,
                               double *out, double *policy_out, int32_t *n_iter, int32_t *status,
                               double timeout_s, int chunk, void *work, size_t work_bytes, void *stream);

/* ------------------------------------------------------------------------- *
 * dense feature products on the path: reward = features . theta (maxent.py:244)
 * and grad = e_features - features^T . svf (:248).  features [S][F] row-major.
 * ------------------------------------------------------------------------- */
int irlb200_features_dot(const double *features, int S, int F, const double *theta,
                         double *reward, void *stream);
int irlb200_features_grad(const double *features, int S, int F, const double *svf,
                          const double *e_features, double *grad, void *stream);

/* ------------------------------------------------------------------------- *
 * batched DENSE path (BASELINE north_star (4), configs[3]): B reward candidates / policies that share one
 * dense p_transition.  The reference's sweep is then literally `p[a].dot(X)` for a matrix X [S x B]
 * (maxent.py:155, :329; `p[a].T.dot(p_action[:, a] * d)`, :109): an FP64 contraction on the tensor cores
 * (mma.sync m8n8k4 f64) with the elementwise part of the reference as epilogue, every candidate stopped at
 * its own sweep.  For tables with K ~ S successors per state; sparse worlds belong to the ELL kernels above.
 *   irlb200_dense_pack: P [S][S'][A] (A innermost, gridworld.py:124-142) -> packed
 *       [A*S][S] per-action rows | [S][S] action-summed | [S][A*S] transposed   (irlb200_dense_pack_doubles doubles)
 *   work: caller-owned device buffer of irlb200_dense_batch_work_bytes(S, A, B) bytes, private to the call.
 *   All [B][...] arrays are candidate-major.  These calls synchronise `stream` (the host polls the number
 *   of live candidates every few sweeps).
 *   backward: local_action_probabilities (maxent.py:119-159), n_sweeps partition sweeps (reference: 2 S), exact
 *       power-of-two rescale per candidate; iterate [B][S] scratch; policy [B][S][A] out
 *   succ op 1: local_causal_action_probabilities (:279-341), phi [S], value [B][S] out, policy [B][S][A] out
 *        op 2: value_iteration (solver.py:9-52; vi_mean: :55-104), value [B][S] out, policy unused
 *   svf: expected_svf_from_policy (:63-114), policy [B][S][A] in, svf [B][S] out, optional grad = e_features - svf
 * ------------------------------------------------------------------------- */
size_t irlb200_dense_pack_doubles(int S, int A);
size_t irlb200_dense_batch_work_bytes(int S, int A, int B);
int irlb200_dense_pack(const double *P, int S, int A, double *packed, void *stream);
int irlb200_dense_batch_backward(const double *packed, int S, int A, int B, const double *reward,
                                 const uint8_t *terminal_mask, int n_sweeps, double *policy, double *iterate,
                                 void *work, size_t work_bytes, void *stream);
int irlb200_dense_batch_succ(int op, const double *packed, int S, int A, int B, const double *reward,
                             const double *phi, double discount, double eps, int max_sweeps, int vi_mean,
                             double *value, double *policy, int32_t *n_iter, int32_t *status, void *work,
                             size_t work_bytes, void *stream);
int irlb200_dense_batch_svf(const double *packed, int S, int A, int B, const double *p_initial, int p0_shared,
                            const uint8_t *terminal_mask, const double *policy, double eps, int max_sweeps,
                            double *svf, const double *e_features, int ef_shared, double *grad,
                            int32_t *n_iter, int32_t *status, void *work, size_t work_bytes, void *stream);

/* ------------------------------------------------------------------------- *
 * expert demonstrations on the device
 *   replaces trajectory.generate_trajectories / generate_trajectory, trajectory.py:52-128, whose
 *   per-step `np.random.choice` over the dense row p_transition[s, :, a] is O(S) and needs a table
 *   that does not exist for large worlds, and folds in the statistics of maxent.py:15-60
 *   (feature_expectation_from_trajectories for one-hot features, initial_probabilities_from_trajectories).
 *   n_traj independent rollouts of the stochastic policy [S][A] (a deterministic policy is a one-hot
 *   row) through the successor table of ONE world, each from a start state drawn from start_cdf [S]
 *   (inclusive running sum of the start distribution) until a state with terminal_mask != 0 or
 *   max_len transitions (then *n_truncated is incremented; the reference would keep going).
 *   Counter-based generator (Philox4x32-10; key = seed, counter = (step, trajectory)): trajectory i
 *   depends on (seed, i) only.  numpy's global Mersenne-Twister stream is NOT reproduced -- seeded
 *   runs of the reference's own sampler are reproduced by the host-side trajectory.py of this package.
 *   states  [n_traj][max_len+1] int32 out or NULL   (row i: lengths[i]+1 valid entries)
 *   actions [n_traj][max_len]   int32 out or NULL
 *   lengths [n_traj] int32 out (transitions per trajectory)
 *   visit_counts [S] f64, start_counts [S] f64: ACCUMULATED into (zero them first), or NULL
 *   n_truncated  [1] int32, accumulated into
 * ------------------------------------------------------------------------- */
int irlb200_sample_trajectories(const irlb200_tables *t, const double *policy, const double *start_cdf,
                                const uint8_t *terminal_mask, int n_traj, int max_len, uint64_t seed,
                                int32_t *states, int32_t *actions, int32_t *lengths,
                                double *visit_counts, double *start_counts, int32_t *n_truncated,
                                void *stream);

#ifdef __cplusplus
}
#endif
#endif /* IRL_MAXENT_B200_H */
