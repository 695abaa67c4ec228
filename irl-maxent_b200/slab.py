"""
slab -- one huge grid MDP sharded by state-row slabs over the ranks of a
`torch.distributed` process group (BASELINE configs[4]).

Rank k owns the grid rows [r0_k, r1_k), i.e. the contiguous states
[lo, hi) = [r0_k * n, r1_k * n), and only their table rows.  Successor and
predecessor offsets of a grid world are {-n, -1, 0, +1, +n}, so a sweep over
the owned states needs exactly one ghost row from each neighbouring rank.

Per sweep:   local sweep kernel (irlb200_slab_sweep, one launch)
             -> ghost-row exchange with <= 2 neighbours (send/recv)
Per chunk:   ONE all-reduce(MAX) of the chunk's per-sweep votes and ONE host
             read.  The reference stops at the first sweep whose max |diff| <= eps
             (maxent.py:108,326; solver.py:40); to return exactly that iterate
             and that sweep count, the iterate at the start of the chunk is
             snapshotted and, once the stopping sweep is known, replayed up to it.

This is the NCCL baseline of the mode (collectives between launches).  The
arithmetic backend is injectable so that the exchange / vote / replay protocol
is covered by world_size-2 `gloo` tests on CPU (tests/test_slab_gloo.py supplies
a numpy sweep there); the product backend below is CUDA-only, no CPU fallback.
"""

import numpy as np

ST_CONVERGED, ST_NONFINITE, ST_MAXSWEEPS, ST_ABORTED = 0, 1, 2, 3
OP_SOFT_VI, OP_VI, OP_SVF = 1, 2, 3
NEG_HUGE = -1e200           # reference: maxent.py:323


def row_partition(n_rows, world):
    """Contiguous, near-equal split of the grid rows: list of (r0, r1) per rank."""
    base, extra = divmod(n_rows, world)
    out, r = [], 0
    for k in range(world):
        c = base + (1 if k < extra else 0)
        out.append((r, r + c))
        r += c
    return out


class CudaBackend:
    """Product backend: the sm_100a kernels behind the C ABI."""

    def __init__(self):
        import _irlb200 as E
        self.E = E
        self.torch = E.require_cuda()
        self.device = E._dev()

    def local_tables(self, size, p_slip, icy, lo, cnt, slots=4):
        E, torch = self.E, self.torch
        K, A = slots, 4                 # slabs are streamed every sweep: the compact 4-slot form
        t = dict(A=A, K=K,
                 succ_idx=torch.empty((K, cnt), dtype=torch.int32, device=self.device),
                 succ_p=torch.empty((A, K, cnt), dtype=torch.float64, device=self.device),
                 pred_idx=torch.empty((K, cnt), dtype=torch.int32, device=self.device),
                 pred_p=torch.empty((A, K, cnt), dtype=torch.float64, device=self.device))
        E._check(E._lib.irlb200_gridworld_tables_range_k(size, 1 if icy else 0, float(p_slip), lo, cnt, K,
                                                         E._ptr(t["succ_idx"]), E._ptr(t["succ_p"]),
                                                         E._ptr(t["pred_idx"]), E._ptr(t["pred_p"]), E._stream()))
        return t

    def sweep(self, op, lo, cnt, A, K, idx, p, c0, c1, discount, eps, vi_mean, x_in, x_out, vote, policy):
        E = self.E
        with E._timed("slab_sweep"):
            E._check(E._lib.irlb200_slab_sweep(op, lo, cnt, A, K, E._ptr(idx), E._ptr(p), E._ptr(c0), E._ptr(c1),
                                               float(discount), float(eps), int(vi_mean), E._ptr(x_in), E._ptr(x_out),
                                               E._ptr(vote), E._ptr(policy), E._stream()))

    def weights(self, cnt, A, K, pred_idx, pred_p, policy_full, mask_full, W):
        E = self.E
        with E._timed("slab_weights"):
            E._check(E._lib.irlb200_slab_weights(cnt, A, K, E._ptr(pred_idx), E._ptr(pred_p), E._ptr(policy_full),
                                                 E._ptr(mask_full), E._ptr(W), E._stream()))


class SlabGrid:
    """One GridWorld / IcyGridWorld of side `size`, sharded by rows over `group`."""

    def __init__(self, size, p_slip=0.2, icy=True, group=None, backend=None, chunk=32):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if size < self.world:
            raise ValueError("more ranks than grid rows")
        self.size, self.n_states = size, size * size
        self.backend = backend or CudaBackend()
        self.device = self.backend.device
        self.chunk = int(chunk)
        r0, r1 = row_partition(size, self.world)[self.rank]
        self.lo, self.hi = r0 * size, r1 * size
        self.cnt, self.halo = self.hi - self.lo, size
        self.tables = self.backend.local_tables(size, p_slip, icy, self.lo, self.cnt)
        self.A, self.K = self.tables["A"], self.tables["K"]
        self.last_n_iter, self.last_status = None, None
        self.n_exchanges = 0

    # -- communication ---------------------------------------------------------------
    def exchange(self, x, width=1):
        """Fill the ghost rows of the full-length vector `x` (`width` values per state)."""
        if self.world == 1:
            return
        dist, h = self.dist, self.halo * width
        lo, hi = self.lo * width, self.hi * width
        flat = x.view(-1)
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, flat[lo:lo + h], self._peer(self.rank - 1), self.group))
            ops.append(dist.P2POp(dist.irecv, flat[lo - h:lo], self._peer(self.rank - 1), self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, flat[hi - h:hi], self._peer(self.rank + 1), self.group))
            ops.append(dist.P2POp(dist.irecv, flat[hi:hi + h], self._peer(self.rank + 1), self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.n_exchanges += 1

    def _peer(self, group_rank):
        return self.dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    def _reduce_votes(self, votes):
        if self.world > 1:
            self.dist.all_reduce(votes, op=self.dist.ReduceOp.MAX, group=self.group)
        return votes.cpu().numpy()

    def local(self, full):
        """Owned slice of a full-length host/device vector."""
        return full[self.lo:self.hi]

    def _dev(self, x, dtype=None):
        torch = self.torch
        dtype = dtype or torch.float64
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device)

    # -- the shared fixed-point driver --------------------------------------------------
    def _iterate(self, op, idx, p, c0, c1, x0_local, discount, eps, max_sweeps, vi_mean=0, policy_out=None):
        torch = self.torch
        S, lo, hi, be = self.n_states, self.lo, self.hi, self.backend
        x = [torch.zeros(S, dtype=torch.float64, device=self.device) for _ in range(2)]
        x[0][lo:hi] = x0_local
        self.exchange(x[0])                                    # ghost rows of the initial iterate
        limit = int(max_sweeps) if max_sweeps and max_sweeps > 0 else None
        n, status = 0, ST_CONVERGED
        null = None

        def run(b, vote, pol):
            be.sweep(op, lo, self.cnt, self.A, self.K, idx, p, c0, c1, discount, eps, vi_mean,
                     x[b], x[b ^ 1], vote, pol)
            self.exchange(x[b ^ 1])

        while True:
            k = self.chunk if limit is None else min(self.chunk, limit - n)
            if k <= 0:
                status = ST_MAXSWEEPS
                break
            snapshot = x[n & 1].clone()
            votes = torch.zeros((k, 2), dtype=torch.int32, device=self.device)
            for i in range(k):
                run((n + i) & 1, votes[i], null)
            v = self._reduce_votes(votes)                     # one collective + one host read per chunk
            stop = None
            for i in range(k):
                if v[i, 1] != 0 or v[i, 0] == 0:               # NaN ends the loop; so does delta <= eps
                    stop = i
                    break
            if stop is None:
                n += k
                continue
            # replay exactly up to the reference's last sweep from the snapshot
            x[n & 1].copy_(snapshot)
            scratch = torch.zeros(2, dtype=torch.int32, device=self.device)
            for i in range(stop + 1):
                run((n + i) & 1, scratch, policy_out if i == stop else null)
            n += stop + 1
            status = ST_NONFINITE if v[stop, 1] != 0 else ST_CONVERGED
            break
        if status == ST_MAXSWEEPS and policy_out is not None and n > 0:
            # guard hit: emit the policy of the last executed sweep
            b = (n - 1) & 1
            scratch = torch.zeros(2, dtype=torch.int32, device=self.device)
            be.sweep(op, lo, self.cnt, self.A, self.K, idx, p, c0, c1, discount, eps, vi_mean,
                     x[b], x[b ^ 1], scratch, policy_out)
        self.last_n_iter, self.last_status = n, status
        return x[n & 1]

    # -- public fixed points (arguments are the OWNED slices, length cnt) ------------------
    def soft_vi(self, reward_local, phi_local, discount, eps=1e-5, max_sweeps=None):
        """local_causal_action_probabilities (maxent.py:279-341) on the slab: returns
        (policy [cnt, A], value [cnt])."""
        torch, t = self.torch, self.tables
        r, phi = self._dev(reward_local), self._dev(phi_local)
        pol = torch.empty((self.cnt, self.A), dtype=torch.float64, device=self.device)
        x0 = torch.full((self.cnt,), NEG_HUGE, dtype=torch.float64, device=self.device)
        v = self._iterate(OP_SOFT_VI, t["succ_idx"], t["succ_p"], r, phi, x0, discount, eps, max_sweeps,
                          policy_out=pol)
        return pol, v[self.lo:self.hi]

    def value_iteration(self, reward_local, discount, eps=1e-3, max_sweeps=None, mean=False):
        """solver.value_iteration (solver.py:9-52) on the slab: returns value [cnt]."""
        torch, t = self.torch, self.tables
        r = self._dev(reward_local)
        x0 = torch.zeros(self.cnt, dtype=torch.float64, device=self.device)
        v = self._iterate(OP_VI, t["succ_idx"], t["succ_p"], r, None, x0, discount, eps, max_sweeps,
                          vi_mean=1 if mean else 0)
        return v[self.lo:self.hi]

    def svf(self, p_initial_local, terminal, policy_local, eps=1e-5, max_sweeps=None):
        """expected_svf_from_policy (maxent.py:63-114) on the slab: returns d [cnt].
        `terminal`: global state indices."""
        torch, t = self.torch, self.tables
        S = self.n_states
        pol_full = torch.zeros((S, self.A), dtype=torch.float64, device=self.device)
        pol_full[self.lo:self.hi] = self._dev(policy_local)
        self.exchange(pol_full, width=self.A)                 # ghost rows of the policy, once
        mask = np.zeros(S, dtype=np.uint8)
        mask[np.asarray(list(terminal), dtype=np.int64)] = 1
        mask_d = torch.as_tensor(mask).to(self.device)
        W = torch.empty((self.K, self.cnt), dtype=torch.float64, device=self.device)
        self.backend.weights(self.cnt, self.A, self.K, t["pred_idx"], t["pred_p"], pol_full, mask_d, W)
        p0 = self._dev(p_initial_local)
        x0 = torch.zeros(self.cnt, dtype=torch.float64, device=self.device)
        d = self._iterate(OP_SVF, t["pred_idx"], W, p0, None, x0, 0.0, eps, max_sweeps)
        return d[self.lo:self.hi]

    def gather(self, local_vec):
        """All ranks' owned slices concatenated (full-length vector on every rank)."""
        if self.world == 1:
            return local_vec.clone()
        dist, torch = self.dist, self.torch
        sizes = [(r1 - r0) * self.size for r0, r1 in row_partition(self.size, self.world)]
        shape_tail = tuple(local_vec.shape[1:])
        m = max(sizes)                                         # equal-sized pieces for every backend
        mine = torch.zeros((m,) + shape_tail, dtype=local_vec.dtype, device=local_vec.device)
        mine[:local_vec.shape[0]] = local_vec
        parts = [torch.empty_like(mine) for _ in sizes]
        dist.all_gather(parts, mine, group=self.group)
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


class PeerSlabGrid(SlabGrid):
    """Slab mode with the halo exchange inside the kernel: one persistent cooperative launch per
    rank and fixed point (`irlb200_slab_persistent`), boundary rows pushed into the neighbours'
    ghost rows with NVLink peer stores, stop rule all-reduced through peer-mapped flag words.
    torch.distributed is used for plumbing only: exchanging the CUDA IPC handles once, one
    ghost-row exchange of the policy before the forward pass, and a process-group barrier
    around each launch.  CUDA only."""

    def __init__(self, size, p_slip=0.2, icy=True, group=None, timeout_s=20.0, overlap=True, flow=True,
                 chunk_sweeps=0):
        super().__init__(size, p_slip, icy, group=group, backend=CudaBackend())
        import os
        # boundary-first kernel: halo exchange overlaps the interior sweep (IRLB200_SLAB_OVERLAP=0/1 overrides)
        self.overlap = int(os.environ.get("IRLB200_SLAB_OVERLAP", "1" if overlap else "0"))
        # dataflow kernel (csrc/slab_flow.cu): neighbour flags instead of a barrier per sweep; the default.
        # IRLB200_SLAB_FLOW=0 selects the barrier-per-sweep kernels (cross-check variant)
        self.flow = int(os.environ.get("IRLB200_SLAB_FLOW", "1" if flow else "0"))
        self.chunk_sweeps = int(chunk_sweeps)
        self._work = None
        # dictionary-coded successor probabilities for the soft-VI / VI sweeps of the dataflow kernel: a grid world's
        # table holds a handful of distinct values, so the sweeps read one byte per entry instead of eight
        self._code, self._dict = None, None
        if self.flow and os.environ.get("IRLB200_FLOW_CODED", "1") != "0":
            torch = self.torch
            sp = self.tables["succ_p"]
            vals = torch.unique(sp)
            if 0 < vals.numel() <= 256:
                self._code = torch.searchsorted(vals, sp.reshape(-1)).to(torch.uint8).reshape(sp.shape).contiguous()
                self._dict = torch.zeros(256, dtype=torch.float64, device=sp.device)
                self._dict[:vals.numel()] = vals
                assert bool((self._dict[self._code.long()] == sp).all()), "dictionary coding must be exact"
        import ctypes
        E = self.backend.E
        self.E, self.ct = E, ctypes
        self.timeout_s = float(timeout_s)
        nbytes = E._lib.irlb200_slab_flow_block_bytes(self.n_states, self.halo)
        own = ctypes.c_void_p()
        E._check(E._lib.irlb200_peer_alloc(nbytes, ctypes.byref(own)))
        self._own, self._nbytes = own, nbytes
        self._imported = []
        ptrs = [None] * self.world
        ptrs[self.rank] = own.value
        if self.world > 1:
            handle = ctypes.create_string_buffer(64)
            E._check(E._lib.irlb200_ipc_export(own, handle))
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, bytes(handle.raw), group=self.group)
            for r, h in enumerate(gathered):
                if r == self.rank:
                    continue
                p = ctypes.c_void_p()
                E._check(E._lib.irlb200_ipc_import(ctypes.create_string_buffer(h, 64), ctypes.byref(p)))
                ptrs[r] = p.value
                self._imported.append(p)
        self._blocks = (ctypes.c_void_p * self.world)(*ptrs)

    def close(self):
        E = self.E
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        for p in self._imported:
            E._lib.irlb200_ipc_close(p)
        self._imported = []
        if self._own is not None:
            E._lib.irlb200_peer_free(self._own)
            self._own = None

    def _fence(self):
        """All ranks idle and every header zero before anybody launches."""
        torch, E = self.torch, self.E
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        E._check(E._lib.irlb200_slab_flow_reset(self._own, self.n_states, self.halo, E._stream()))
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def _run(self, op, idx, p, c0, c1, policy_in, mask, w_scratch, discount, eps, max_sweeps, vi_mean, policy_out):
        torch, E = self.torch, self.E
        out = torch.empty(self.cnt, dtype=torch.float64, device=self.device)
        n_iter = torch.zeros(1, dtype=torch.int32, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._fence()
        ms = E.DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
        with E._timed("slab_persistent"):
            if self.flow:
                if self._work is None:
                    nbytes = int(E._lib.irlb200_slab_flow_work_bytes(self.cnt))
                    self._work = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                coded = op != OP_SVF and self._code is not None
                E._check(E._lib.irlb200_slab_flow_coded(
                    op, self.rank, self.world, self._blocks, self.n_states, self.lo, self.cnt, self.halo, self.A,
                    self.K, E._ptr(idx), E._ptr(p), E._ptr(self._code if coded else None),
                    E._ptr(self._dict if coded else None), E._ptr(c0), E._ptr(c1), E._ptr(policy_in), E._ptr(mask),
                    E._ptr(w_scratch), float(discount), float(eps), ms, int(vi_mean), E._ptr(out), E._ptr(policy_out),
                    E._ptr(n_iter), E._ptr(status), self.timeout_s, self.chunk_sweeps, E._ptr(self._work),
                    self._work.numel(), E._stream()))
            else:
                E._check(E._lib.irlb200_slab_persistent(
                    op, self.rank, self.world, self._blocks, self.n_states, self.lo, self.cnt, self.halo, self.A,
                    self.K, E._ptr(idx), E._ptr(p), E._ptr(c0), E._ptr(c1), E._ptr(policy_in), E._ptr(mask),
                    E._ptr(w_scratch), float(discount), float(eps), ms, int(vi_mean), E._ptr(out), E._ptr(policy_out),
                    E._ptr(n_iter), E._ptr(status), self.timeout_s, int(self.overlap), E._stream()))
        torch.cuda.synchronize()
        self.last_n_iter, self.last_status = int(n_iter.item()), int(status.item())
        if self.last_status == ST_ABORTED:
            raise RuntimeError("slab kernel aborted: a peer rank did not arrive within %.0f s" % self.timeout_s)
        return out

    def soft_vi(self, reward_local, phi_local, discount, eps=1e-5, max_sweeps=None):
        torch, t = self.torch, self.tables
        r, phi = self._dev(reward_local), self._dev(phi_local)
        pol = torch.empty((self.cnt, self.A), dtype=torch.float64, device=self.device)
        v = self._run(OP_SOFT_VI, t["succ_idx"], t["succ_p"], r, phi, None, None, None, discount, eps, max_sweeps, 0, pol)
        return pol, v

    def value_iteration(self, reward_local, discount, eps=1e-3, max_sweeps=None, mean=False):
        t = self.tables
        r = self._dev(reward_local)
        return self._run(OP_VI, t["succ_idx"], t["succ_p"], r, None, None, None, None, discount, eps, max_sweeps,
                         1 if mean else 0, None)

    def svf(self, p_initial_local, terminal, policy_local, eps=1e-5, max_sweeps=None):
        torch, t = self.torch, self.tables
        S = self.n_states
        pol_full = torch.zeros((S, self.A), dtype=torch.float64, device=self.device)
        pol_full[self.lo:self.hi] = self._dev(policy_local)
        self.exchange(pol_full, width=self.A)                 # ghost rows of the policy, once per pass
        mask = np.zeros(S, dtype=np.uint8)
        mask[np.asarray(list(terminal), dtype=np.int64)] = 1
        mask_d = torch.as_tensor(mask).to(self.device)
        W = torch.empty((self.K, self.cnt), dtype=torch.float64, device=self.device)
        p0 = self._dev(p_initial_local)
        return self._run(OP_SVF, t["pred_idx"], t["pred_p"], p0, None, pol_full, mask_d, W, 0.0, eps, max_sweeps, 0, None)
