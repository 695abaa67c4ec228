"""
_irlb200 -- host-side engine: ctypes binding of libirlmaxent_b200.so (the C ABI
declared in include/irl_maxent_b200.h) plus the device-resident table handle.

PyTorch is used for plumbing only (device memory, streams, host<->device
copies); every arithmetic step of the hot path runs in the hand-written
sm_100a kernels behind the C ABI.  There is NO CPU fallback: if the shared
library is missing or no CUDA device is visible, every compute call raises.
"""

import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libirlmaxent_b200.so")

MODE_AUTO, MODE_CTA, MODE_CLUSTER, MODE_GRID = 0, 1, 2, 3
ST_CONVERGED, ST_NONFINITE, ST_MAXSWEEPS, ST_ABORTED = 0, 1, 2, 3

# guard against inputs for which the reference would loop forever
# (non-absorbing policies, maxent.py:108); 0 disables the guard
DEFAULT_MAX_SWEEPS = int(os.environ.get("IRLB200_MAX_SWEEPS", "200000000"))


class EngineError(RuntimeError):
    pass


class _CTables(ctypes.Structure):
    _fields_ = [("S", ctypes.c_int32), ("A", ctypes.c_int32),
                ("Ks", ctypes.c_int32), ("Kp", ctypes.c_int32),
                ("succ_idx", ctypes.c_void_p), ("succ_p", ctypes.c_void_p),
                ("pred_idx", ctypes.c_void_p), ("pred_p", ctypes.c_void_p),
                ("shared", ctypes.c_int32), ("stencil_n", ctypes.c_int32)]


_lib = None

_vp, _i, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
_tp = ctypes.POINTER(_CTables)

# name -> argtypes; mirrors include/irl_maxent_b200.h one to one
SIGNATURES = {
    "irlb200_version": ([], _i),
    "irlb200_last_error": ([], ctypes.c_char_p),
    "irlb200_device_count": ([], _i),
    "irlb200_max_states_cta": ([], _i),
    "irlb200_max_states_cluster": ([], _i),
    "irlb200_dense_count": ([_vp, _i, _i, _vp, _vp, _vp, _vp], _i),
    "irlb200_dense_fill": ([_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_gridworld_tables": ([_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_gridworld_tables_k": ([_i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_gridworld_tables_range_k": ([_i, _i, _d, _i, _i, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_gridworld_dense": ([_i, _i, _d, _vp, _vp], _i),
    "irlb200_gridworld_tables_range": ([_i, _i, _d, _i, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_slab_sweep": ([_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _d, _d, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_slab_weights": ([_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_peer_alloc": ([ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)], _i),
    "irlb200_peer_free": ([_vp], _i),
    "irlb200_ipc_export": ([_vp, ctypes.c_char_p], _i),
    "irlb200_ipc_import": ([ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)], _i),
    "irlb200_ipc_close": ([_vp], _i),
    "irlb200_slab_block_bytes": ([_i], ctypes.c_size_t),
    "irlb200_slab_reset": ([_vp, _vp], _i),
    "irlb200_slab_persistent": ([_i, _i, _i, ctypes.POINTER(ctypes.c_void_p), _i, _i, _i, _i, _i, _i, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp, _vp, _d, _i, _vp], _i),
    "irlb200_slab_flow_work_bytes": ([_i], ctypes.c_size_t),
    "irlb200_slab_flow_block_bytes": ([_i, _i], ctypes.c_size_t),
    "irlb200_slab_flow_reset": ([_vp, _i, _i, _vp], _i),
    "irlb200_slab_flow": ([_i, _i, _i, ctypes.POINTER(ctypes.c_void_p), _i, _i, _i, _i, _i, _i, _vp, _vp, _vp,
                           _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp, _vp, _d, _i, _vp, ctypes.c_size_t,
                           _vp], _i),
    "irlb200_slab_flow_coded": ([_i, _i, _i, ctypes.POINTER(ctypes.c_void_p), _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp, _vp, _d, _i, _vp,
                                 ctypes.c_size_t, _vp], _i),
    "irlb200_backward": ([_tp, _i, _vp, _vp, _i, _i, _vp, _i, _vp], _i),
    "irlb200_soft_vi": ([_tp, _i, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "irlb200_value_iteration": ([_tp, _i, _vp, _d, _d, _i, _i, _vp, _vp, _vp, _i, _vp], _i),
    "irlb200_svf": ([_tp, _i, _vp, _i, _vp, _i, _vp, _d, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp], _i),
    "irlb200_svf_ordered": ([_tp, _i, _vp, _i, _vp, _i, _vp, _d, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp], _i),
    "irlb200_expected_svf": ([_tp, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _d, _d, _d, _i,
                              _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_dense_pack_doubles": ([_i, _i], ctypes.c_size_t),
    "irlb200_dense_batch_work_bytes": ([_i, _i, _i], ctypes.c_size_t),
    "irlb200_dense_pack": ([_vp, _i, _i, _vp, _vp], _i),
    "irlb200_dense_batch_backward": ([_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, ctypes.c_size_t, _vp], _i),
    "irlb200_dense_batch_succ": ([_i, _vp, _i, _i, _i, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                  ctypes.c_size_t, _vp], _i),
    "irlb200_dense_batch_svf": ([_vp, _i, _i, _i, _vp, _i, _vp, _vp, _d, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp,
                                 ctypes.c_size_t, _vp], _i),
    "irlb200_irl_small": ([_tp, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _d, _d, _d, _i, _i, _vp, _i, _i, _d,
                           _vp, _vp, _vp, _vp], _i),
    "irlb200_sample_trajectories": ([_tp, _vp, _vp, _vp, _i, _i, ctypes.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "irlb200_features_dot": ([_vp, _i, _i, _vp, _vp, _vp], _i),
    "irlb200_features_grad": ([_vp, _i, _i, _vp, _vp, _vp, _vp], _i),
}


def load_library(path=None):
    """dlopen the C-ABI library and declare every prototype.  No compute happens here."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EngineError(
            "native library %s is missing: build it with `python irl-maxent_b200/build_native.py` "
            "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)         # AttributeError if the symbol is not exported
        fn.argtypes, fn.restype = argtypes, restype
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise EngineError("libirlmaxent_b200: error %d: %s" % (rc, _lib.irlb200_last_error().decode()))


def _torch():
    import torch
    return torch


def require_cuda():
    torch = _torch()
    lib = load_library()
    if not torch.cuda.is_available() or lib.irlb200_device_count() <= 0:
        raise EngineError("no CUDA device: the B200 engine has no CPU fallback")
    return torch


def _stream():
    return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _dev():
    return _torch().device("cuda", _torch().cuda.current_device())


def to_device(x, dtype=None):
    """numpy / list / tensor -> contiguous CUDA tensor (float64 unless dtype given)."""
    torch = _torch()
    dtype = dtype or torch.float64
    if isinstance(x, torch.Tensor):
        return x.to(device=_dev(), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(_dev())


def is_tensor(x):
    try:
        import torch
        return isinstance(x, torch.Tensor)
    except ImportError:                                         # pragma: no cover
        return False


# ---------------------------------------------------------------------------
# tables
# ---------------------------------------------------------------------------

class Tables:
    """Device-resident state-merged ELL tables of B (or one shared) MDP(s).

    Layout as in include/irl_maxent_b200.h: succ_idx [Bt][Ks][S] int32,
    succ_p [Bt][A][Ks][S] f64, pred_idx [Bt][Kp][S], pred_p [Bt][A][Kp][S].
    """

    def __init__(self, S, A, Ks, Kp, succ_idx, succ_p, pred_idx, pred_p, n_tables=1, stencil_n=0):
        self.S, self.A, self.Ks, self.Kp = int(S), int(A), int(Ks), int(Kp)
        self.succ_idx, self.succ_p, self.pred_idx, self.pred_p = succ_idx, succ_p, pred_idx, pred_p
        self.n_tables = int(n_tables)
        # > 0: S = n*n and all links stay within {s-n, s-1, s, s+1, s+n} without wrapping a row end
        self.stencil_n = int(stencil_n)

    # the reference's `n_states, _, n_actions = p_transition.shape` keeps working on a handle
    @property
    def shape(self):
        return (self.S, self.S, self.A)

    def c_struct(self, shared):
        return _CTables(self.S, self.A, self.Ks, self.Kp, self.succ_idx.data_ptr(), self.succ_p.data_ptr(),
                        self.pred_idx.data_ptr(), self.pred_p.data_ptr(), 1 if shared else 0, self.stencil_n)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.succ_idx, self.succ_p, self.pred_idx, self.pred_p))

    def select(self, b):
        """Handle on the tables of world b of a batch (views, no copy)."""
        return Tables(self.S, self.A, self.Ks, self.Kp, self.succ_idx[b:b + 1], self.succ_p[b:b + 1],
                      self.pred_idx[b:b + 1], self.pred_p[b:b + 1], 1, self.stencil_n)

    def detect_stencil(self):
        """Set stencil_n if the tables have the 5-point grid structure (torch ops, once per table)."""
        torch = _torch()
        n = int(round(self.S ** 0.5))
        self.stencil_n = 0
        if n * n != self.S or n < 2 or self.A != 4 or self.Ks != 5 or self.Kp != 5:
            return 0
        s = torch.arange(self.S, device=self.succ_idx.device, dtype=torch.int64)
        x = s % n
        for idx, p in ((self.succ_idx, self.succ_p), (self.pred_idx, self.pred_p)):
            off = idx.to(torch.int64) - s                      # [Bt, K, S]
            used = (p != 0).any(dim=1)                         # [Bt, K, S]
            ok = (off == 0) | (off == n) | (off == -n) | ((off == 1) & (x < n - 1)) | ((off == -1) & (x > 0))
            if not bool((ok | ~used).all()):
                return 0
        self.stencil_n = n
        return n


def _round_slots(A, k):
    # every 4-action table with <= 5 neighbours is padded to the register-resident shape
    k = max(int(k), 1)
    return 5 if (A == 4 and k <= 5) else k


def compress_dense(p_transition):
    """Kernel (1): dense p_transition[S,S',A] -> Tables (one-time)."""
    torch = require_cuda()
    P = to_device(p_transition)
    if P.dim() != 3 or P.shape[0] != P.shape[1]:
        raise EngineError("p_transition must have shape [S, S, A]")
    S, A = int(P.shape[0]), int(P.shape[2])
    dev = P.device
    succ_cnt = torch.empty(S, dtype=torch.int32, device=dev)
    pred_cnt = torch.empty(S, dtype=torch.int32, device=dev)
    kmax = torch.empty(2, dtype=torch.int32, device=dev)
    _check(_lib.irlb200_dense_count(_ptr(P), S, A, _ptr(succ_cnt), _ptr(pred_cnt), _ptr(kmax), _stream()))
    ks, kp = (int(v) for v in kmax.tolist())            # host sync: sizes are needed to allocate
    Ks, Kp = _round_slots(A, ks), _round_slots(A, kp)
    succ_idx = torch.empty((1, Ks, S), dtype=torch.int32, device=dev)
    succ_p = torch.empty((1, A, Ks, S), dtype=torch.float64, device=dev)
    pred_idx = torch.empty((1, Kp, S), dtype=torch.int32, device=dev)
    pred_p = torch.empty((1, A, Kp, S), dtype=torch.float64, device=dev)
    _check(_lib.irlb200_dense_fill(_ptr(P), S, A, Ks, Kp, _ptr(succ_idx), _ptr(succ_p), _ptr(pred_idx),
                                   _ptr(pred_p), _ptr(pred_cnt), _stream()))
    t = Tables(S, A, Ks, Kp, succ_idx, succ_p, pred_idx, pred_p, 1)
    t.k_discovered = (ks, kp)
    t.detect_stencil()
    return t


def gridworld_tables(size, p_slip=None, icy=True, slots=5):
    """Tables of GridWorld / IcyGridWorld straight from (size, p_slip).
    `p_slip` may be a scalar or a length-B sequence (B worlds).
    `slots`: 5 = one slot per stencil position (register-resident, tiled and cluster kernels);
    4 = compact form for worlds that are streamed from HBM every sweep (15-22 % fewer bytes,
    bitwise the same results)."""
    torch = require_cuda()
    if slots not in (4, 5):
        raise EngineError("slots must be 4 or 5")
    if icy:
        ps = np.atleast_1d(np.asarray(p_slip, dtype=np.float64))
    else:
        ps = np.zeros(1)
    B = len(ps)
    S, A, K = size * size, 4, slots
    dev = _dev()
    ps_d = to_device(ps)
    succ_idx = torch.empty((B, K, S), dtype=torch.int32, device=dev)
    succ_p = torch.empty((B, A, K, S), dtype=torch.float64, device=dev)
    pred_idx = torch.empty((B, K, S), dtype=torch.int32, device=dev)
    pred_p = torch.empty((B, A, K, S), dtype=torch.float64, device=dev)
    _check(_lib.irlb200_gridworld_tables_k(size, 1 if icy else 0, B, _ptr(ps_d), K, _ptr(succ_idx), _ptr(succ_p),
                                           _ptr(pred_idx), _ptr(pred_p), _stream()))
    return Tables(S, A, K, K, succ_idx, succ_p, pred_idx, pred_p, B, stencil_n=size if (size >= 2 and K == 5) else 0)


def gridworld_dense(size, p_slip=0.2, icy=True):
    """Dense [S,S,A] table of a grid world, built on the device (float64 CUDA tensor)."""
    torch = require_cuda()
    S = size * size
    P = torch.empty((S, S, 4), dtype=torch.float64, device=_dev())
    _check(_lib.irlb200_gridworld_dense(size, 1 if icy else 0, float(p_slip), _ptr(P), _stream()))
    return P


# ---- cache: the reference API hands the dense table to every call ------------

_cache = {}
_FULL_HASH_BYTES = 64 << 20        # arrays up to this size are fingerprinted in full on every call


def clear_cache():
    """Forget every compressed table (see `as_tables`)."""
    _cache.clear()


def invalidate(p_transition):
    """Forget the tables compressed from this array / tensor (after editing a very large one in place)."""
    _cache.pop(id(p_transition), None)


def _fingerprint(p):
    """Cheap content fingerprint of a dense table, so that an IN-PLACE edit between two calls is seen
    (the reference re-reads the array on every call).  CUDA / CPU tensors: torch's version counter
    (bumped by every in-place op) + storage address.  numpy arrays: crc32 of the whole buffer up to
    64 MiB (33.5 MB = a 32x32 world: ~15 ms); beyond that crc32 of 4 M evenly strided elements plus the
    corners -- an edit that misses all of them needs `invalidate(p)` / `clear_cache()`."""
    import zlib
    if is_tensor(p):
        return ("t", p._version, p.data_ptr(), tuple(p.shape))
    a = np.asarray(p)
    if a.nbytes <= _FULL_HASH_BYTES:
        return ("n", zlib.crc32(np.ascontiguousarray(a).view(np.uint8).reshape(-1)), a.shape)
    flat = a.reshape(-1) if a.flags.c_contiguous else a.ravel()
    step = max(1, flat.size // (4 << 20))
    sample = np.ascontiguousarray(flat[::step])
    return ("s", zlib.crc32(sample.view(np.uint8)), float(flat[0]), float(flat[-1]), a.shape)


def as_tables(p_transition):
    """Tables for whatever the caller passed as `p_transition` (dense array, CUDA
    tensor or an existing handle).  Dense inputs are compressed once and cached by
    object identity + content fingerprint (`_fingerprint`), because the reference
    API passes the same dense table to every call of the inner loop."""
    if isinstance(p_transition, Tables):
        return p_transition
    key = id(p_transition)
    digest = _fingerprint(p_transition)
    hit = _cache.get(key)
    if hit is not None and hit[1] == digest and hit[2]() is p_transition:
        return hit[0]
    t = compress_dense(p_transition)
    try:
        ref = weakref.ref(p_transition, lambda _r, k=key: _cache.pop(k, None))
    except TypeError:
        return t
    _cache[key] = (t, digest, ref)
    return t


# ---------------------------------------------------------------------------
# helpers shared by the wrappers
# ---------------------------------------------------------------------------

def terminal_mask(terminal, S):
    """uint8 [S] mask from an index collection (device)."""
    torch = _torch()
    m = np.zeros(S, dtype=np.uint8)
    raw = np.asarray(list(terminal))
    if raw.size:
        idx = raw.astype(np.int64)
        # the forward pass needs state indices (reference: maxent.py:99 indexes the table with them); a
        # terminal-reward ARRAY is only meaningful for the causal policy pass (maxent.py:312-317)
        if not np.array_equal(idx, raw) or idx.min() < -S or idx.max() >= S:
            raise EngineError("`terminal` must be a collection of state indices in [0, %d) here" % S)
        m[idx] = 1
    return torch.as_tensor(m).to(_dev())


def terminal_phi(terminal, S):
    """maxent.py:312-317: the array itself iff len(terminal) == S, else 0 / -inf."""
    if is_tensor(terminal):
        if terminal.numel() == S:
            return to_device(terminal)
        terminal = terminal.tolist()
    if len(terminal) == S:
        return to_device(np.array(terminal, dtype=float))
    phi = np.full(S, -np.inf)
    phi[np.asarray(list(terminal), dtype=np.int64)] = 0.0
    return to_device(phi)


def _batch2d(x, S):
    """[S] or [B,S] input -> (contiguous [B,S] device tensor, B)."""
    t = to_device(x)
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.shape[-1] != S:
        raise EngineError("expected trailing dimension %d, got %s" % (S, tuple(t.shape)))
    return t.contiguous(), int(t.shape[0])


def _tables_shared(tables, B):
    if tables.n_tables == 1:
        return True
    if tables.n_tables != B:
        raise EngineError("batch of %d problems needs 1 or %d tables, got %d" % (B, B, tables.n_tables))
    return False


def _maybe_shared(x, S, B, dtype=None):
    """[S] (shared) or [B,S] per-problem vector -> (tensor, shared flag)."""
    t = to_device(x, dtype)
    if t.dim() == 1:
        return t.contiguous(), 1
    if t.shape[0] == 1 and B > 1:
        return t[0].contiguous(), 1
    if t.shape[0] != B:
        raise EngineError("per-problem input has batch %d, expected %d" % (t.shape[0], B))
    return t.contiguous(), 0


class SweepInfo:
    """Iteration counts / stop reasons of the last call (device tensors; reading
    them synchronises)."""

    def __init__(self, n_iter, status):
        self.n_iter, self.status = n_iter, status

    def counts(self):
        return self.n_iter.cpu().numpy()

    def stati(self):
        return self.status.cpu().numpy()


last_info = None

# optional per-launch timing: when `launch_log` is a list, every sweep entry point
# appends (name, start_event, end_event), recorded on the launching stream
launch_log = None
n_launches = 0          # launches of this library's kernels since import (bench.py: gpu_launches)


# NVTX ranges around every entry point (tracing), visible in nsys / ncu --nvtx; IRLB200_NVTX=0 disables
_NVTX = os.environ.get("IRLB200_NVTX", "1") != "0"


class _timed:
    def __init__(self, name, launches=1):
        self.name, self.launches = name, launches

    def __enter__(self):
        global n_launches
        n_launches += self.launches
        if _NVTX:
            _torch().cuda.nvtx.range_push("irlb200:" + self.name)     # one range per fixed point / launch group
        if launch_log is not None:
            torch = _torch()
            self.t0 = torch.cuda.Event(enable_timing=True)
            self.t1 = torch.cuda.Event(enable_timing=True)
            self.t0.record()
        return self

    def __exit__(self, *exc):
        if launch_log is not None:
            self.t1.record()
            launch_log.append((self.name, self.t0, self.t1))
        if _NVTX:
            _torch().cuda.nvtx.range_pop()
        return False


# ---------------------------------------------------------------------------
# entry-point wrappers (device tensors in, device tensors out)
# ---------------------------------------------------------------------------

ELIMIT = -3


def _row(x, b, shared):
    return x if shared else x[b]


def _too_large_for_a_batch(tables, B, mode):
    """A batch of problems that do not fit one CTA's shared memory (S above ~14 500; soft-VI / VI at the C3 size)
    cannot run as one launch: the cooperative-grid and cluster modes take one problem per call.  The wrappers
    then run the batch as B single-problem launches (same results, per-problem counts)."""
    if B <= 1 or mode == MODE_CTA:
        return False
    return (2 * tables.S + 34) * 8 > 227 * 1024


def backward(tables, terminal_mask_t, reward, n_sweeps=None, mode=MODE_AUTO):
    """(2) local_action_probabilities, maxent.py:119-159.  reward [S] or [B,S]."""
    torch = require_cuda()
    S, A = tables.S, tables.A
    r, B = _batch2d(reward, S)
    if _too_large_for_a_batch(tables, B, mode) and not (tables.stencil_n and tables.stencil_n <= 128):
        mask, mshared = _maybe_shared(terminal_mask_t, S, B, torch.uint8)
        return torch.cat([backward(tables if tables.n_tables == 1 else tables.select(b), _row(mask, b, mshared), r[b],
                                   n_sweeps, mode) for b in range(B)], 0)
    mask, mshared = _maybe_shared(terminal_mask_t, S, B, torch.uint8)
    pol = torch.empty((B, S, A), dtype=torch.float64, device=r.device)
    ct = tables.c_struct(_tables_shared(tables, B))
    with _timed("backward"):
        _check(_lib.irlb200_backward(ctypes.byref(ct), B, _ptr(r), _ptr(mask), mshared,
                                     2 * S if n_sweeps is None else int(n_sweeps), _ptr(pol), mode, _stream()))
    return pol


def soft_vi(tables, phi, reward, discount, eps=1e-5, max_sweeps=None, mode=MODE_AUTO, want_value=False):
    """(3) local_causal_action_probabilities, maxent.py:279-341."""
    global last_info
    torch = require_cuda()
    S, A = tables.S, tables.A
    r, B = _batch2d(reward, S)
    ph, pshared = _maybe_shared(phi, S, B)
    if _too_large_for_a_batch(tables, B, mode):
        outs, infos = [], []
        for b in range(B):
            outs.append(soft_vi(tables if tables.n_tables == 1 else tables.select(b), _row(ph, b, pshared), r[b],
                                discount, eps, max_sweeps, mode, want_value))
            infos.append(last_info)
        last_info = SweepInfo(torch.cat([i.n_iter for i in infos]), torch.cat([i.status for i in infos]))
        if want_value:
            return torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0)
        return torch.cat(outs, 0)
    pol = torch.empty((B, S, A), dtype=torch.float64, device=r.device)
    val = torch.empty((B, S), dtype=torch.float64, device=r.device) if want_value else None
    n_iter = torch.empty(B, dtype=torch.int32, device=r.device)      # always written by the kernel
    status = torch.empty(B, dtype=torch.int32, device=r.device)
    ct = tables.c_struct(_tables_shared(tables, B))
    ms = DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
    with _timed("soft_vi"):
        _check(_lib.irlb200_soft_vi(ctypes.byref(ct), B, _ptr(r), _ptr(ph), pshared, float(discount), float(eps),
                                    ms, _ptr(pol), _ptr(val), _ptr(n_iter), _ptr(status), mode, _stream()))
    last_info = SweepInfo(n_iter, status)
    return (pol, val) if want_value else pol


def value_iteration(tables, reward, discount, eps=1e-3, max_sweeps=None, mean=False, mode=MODE_AUTO):
    """solver.value_iteration, solver.py:9-52 (mean=True: :55-104)."""
    global last_info
    torch = require_cuda()
    S = tables.S
    r, B = _batch2d(reward, S)
    if _too_large_for_a_batch(tables, B, mode):
        outs, infos = [], []
        for b in range(B):
            outs.append(value_iteration(tables if tables.n_tables == 1 else tables.select(b), r[b], discount, eps,
                                        max_sweeps, mean, mode))
            infos.append(last_info)
        last_info = SweepInfo(torch.cat([i.n_iter for i in infos]), torch.cat([i.status for i in infos]))
        return torch.cat(outs, 0)
    val = torch.empty((B, S), dtype=torch.float64, device=r.device)
    n_iter = torch.empty(B, dtype=torch.int32, device=r.device)      # always written by the kernel
    status = torch.empty(B, dtype=torch.int32, device=r.device)
    ct = tables.c_struct(_tables_shared(tables, B))
    ms = DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
    with _timed("value_iteration"):
        _check(_lib.irlb200_value_iteration(ctypes.byref(ct), B, _ptr(r), float(discount), float(eps), ms,
                                            1 if mean else 0, _ptr(val), _ptr(n_iter), _ptr(status), mode, _stream()))
    last_info = SweepInfo(n_iter, status)
    return val


# Batches whose forward passes differ in length by several x (IcyGridWorld batch of the bench: mean 45 k,
# max 174 k sweeps) leave SMs idle behind the last long worlds when launched in index order.  The sweep
# counts of one gradient step predict the next one's (omega moves little per step), so every batched
# forward launch stores "longest first" as the launch order of the next launch on the same tables.
# Scheduling only: results do not depend on the order.  IRLB200_LPT=0 disables it.
LPT_MIN_BATCH = 256


def _lpt_enabled():
    return os.environ.get("IRLB200_LPT", "1") != "0"


def svf(tables, p_initial, terminal_mask_t, policy, eps=1e-5, max_sweeps=None, e_features=None,
        mode=MODE_AUTO, order="auto"):
    """(4) expected_svf_from_policy, maxent.py:63-114.  policy [S,A] or [B,S,A].
    With e_features ([S] or [B,S], identity features) also returns grad = e_features - svf.
    `order`: launch-order hint for batches (int32 permutation [B] on the device), None = index order,
    "auto" = longest-first by the sweep counts of the previous batched call on these tables."""
    global last_info
    torch = require_cuda()
    S, A = tables.S, tables.A
    pol = to_device(policy)
    if pol.dim() == 2:
        pol = pol.unsqueeze(0)
    B = int(pol.shape[0])
    if tuple(pol.shape[1:]) != (S, A):
        raise EngineError("policy must have shape [S, A] or [B, S, A]")
    p0, p0shared = _maybe_shared(p_initial, S, B)
    mask, mshared = _maybe_shared(terminal_mask_t, S, B, torch.uint8)
    clusterable = tables.stencil_n and tables.stencil_n <= 128 and tables.stencil_n % 4 == 0 and tables.Kp == 5
    if _too_large_for_a_batch(tables, B, mode) and not clusterable:
        ef_s = _maybe_shared(e_features, S, B) if e_features is not None else (None, 1)
        outs, infos = [], []
        for b in range(B):
            outs.append(svf(tables if tables.n_tables == 1 else tables.select(b), _row(p0, b, p0shared),
                            _row(mask, b, mshared), pol[b], eps, max_sweeps,
                            None if e_features is None else _row(ef_s[0], b, ef_s[1]), mode, None))
            infos.append(last_info)
        last_info = SweepInfo(torch.cat([i.n_iter for i in infos]), torch.cat([i.status for i in infos]))
        if e_features is not None:
            return torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0)
        return torch.cat(outs, 0)
    out = torch.empty((B, S), dtype=torch.float64, device=pol.device)
    grad, ef, efshared = None, None, 1
    if e_features is not None:
        ef, efshared = _maybe_shared(e_features, S, B)
        grad = torch.empty((B, S), dtype=torch.float64, device=pol.device)
    n_iter = torch.empty(B, dtype=torch.int32, device=pol.device)
    status = torch.empty(B, dtype=torch.int32, device=pol.device)
    ct = tables.c_struct(_tables_shared(tables, B))
    ms = DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
    auto = isinstance(order, str)
    if auto:
        hint = getattr(tables, "_svf_order", None)
        order = hint if (hint is not None and hint.numel() == B and hint.device == pol.device and _lpt_enabled()) else None
    elif order is not None:
        order = to_device(order, torch.int32)
        if order.numel() != B:
            raise EngineError("order must be a permutation of the %d problems" % B)
    with _timed("svf"):
        _check(_lib.irlb200_svf_ordered(ctypes.byref(ct), B, _ptr(p0), p0shared, _ptr(mask), mshared, _ptr(pol),
                                        float(eps), ms, _ptr(out), _ptr(ef), efshared, _ptr(grad), _ptr(n_iter),
                                        _ptr(status), mode, _ptr(order), _stream()))
    if auto and B >= LPT_MIN_BATCH and _lpt_enabled():
        # device-side argsort, no host sync; used by the next call on the same tables
        tables._svf_order = torch.argsort(n_iter, descending=True, stable=True).to(torch.int32)
    last_info = SweepInfo(n_iter, status)
    return (out, grad) if grad is not None else out


def expected_svf(tables, p_initial, terminal_mask_t, reward, causal=False, phi=None, discount=0.0,
                 eps_lap=1e-5, eps_svf=1e-5, n_backward=None, max_sweeps=None, e_features=None,
                 want_policy=False, fused=None):
    """compute_expected_svf (maxent.py:162-193) / compute_expected_causal_svf (:344-380).

    fused=True: one launch per batch, policy kept in shared memory (lowest latency).
    fused=False: policy pass and forward pass as two launches with their own
    occupancy-optimal shapes (highest batch throughput).  None: fused for small batches of small worlds.
    Returns (svf, grad or None, policy or None)."""
    global last_info
    torch = require_cuda()
    S, A = tables.S, tables.A
    r, B = _batch2d(reward, S)
    if fused is None:
        # one launch (policy kept in shared memory) wins where launch latency matters: small batches of
        # small worlds.  Mid-size grid worlds are faster through the stencil-tiled / cluster kernels.
        tiled = tables.stencil_n > 0 and tables.stencil_n % 4 == 0
        fused = (B < 64 and (S <= 256 or not tiled)) or (S <= 32 and A == 4 and tables.Ks == 5 and tables.Kp == 5)
    # the fused kernel keeps two iterate buffers and the policy ((2 + A) * S + 34 doubles) in one CTA's
    # shared memory; irlb200_max_states_cta() quotes that limit for A = 4
    if fused and (2 + A) * S > 6 * _lib.irlb200_max_states_cta():
        fused = False
    mask, mshared = _maybe_shared(terminal_mask_t, S, B, torch.uint8)
    if not fused:
        if causal:
            pol = soft_vi(tables, phi, r, discount, eps_lap, max_sweeps)
        else:
            pol = backward(tables, mask, r, n_backward)
        info_a = last_info
        res = svf(tables, p_initial, mask, pol, eps_svf, max_sweeps, e_features)
        if causal:
            last_info = SweepInfo(torch.stack([info_a.n_iter, last_info.n_iter], 1),
                                  torch.stack([info_a.status, last_info.status], 1))
        else:
            nb = torch.full_like(last_info.n_iter, 2 * S if n_backward is None else n_backward)
            last_info = SweepInfo(torch.stack([nb, last_info.n_iter], 1),
                                  torch.stack([torch.zeros_like(last_info.status), last_info.status], 1))
        d, g = res if e_features is not None else (res, None)
        return d, g, (pol if want_policy else None)

    p0, p0shared = _maybe_shared(p_initial, S, B)
    ph = None
    if causal:
        ph, pshared = _maybe_shared(phi, S, B)
        if pshared != mshared:
            raise EngineError("phi and terminal mask must both be shared or both per-problem")
    out = torch.empty((B, S), dtype=torch.float64, device=r.device)
    grad, ef, efshared = None, None, 1
    if e_features is not None:
        ef, efshared = _maybe_shared(e_features, S, B)
        grad = torch.empty((B, S), dtype=torch.float64, device=r.device)
    pol = torch.empty((B, S, A), dtype=torch.float64, device=r.device) if want_policy else None
    n_iter = torch.empty((B, 2), dtype=torch.int32, device=r.device)
    status = torch.empty((B, 2), dtype=torch.int32, device=r.device)
    ct = tables.c_struct(_tables_shared(tables, B))
    ms = DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
    with _timed("expected_svf_fused"):
        _check(_lib.irlb200_expected_svf(
            ctypes.byref(ct), B, 1 if causal else 0, _ptr(r), _ptr(p0), p0shared, _ptr(mask), _ptr(ph), mshared,
            2 * S if n_backward is None else int(n_backward), float(discount), float(eps_lap), float(eps_svf), ms,
            _ptr(out), _ptr(ef), efshared, _ptr(grad), _ptr(pol), _ptr(n_iter), _ptr(status), _stream()))
    last_info = SweepInfo(n_iter, status)
    return out, grad, pol


def sample_trajectories(tables, policy, start, terminal_mask_t, n, seed, max_len=None, store=True):
    """Device rollouts of a stochastic policy [S, A] through the successor table of one world
    (trajectory.py:52-128), with the visit / start-state counts of maxent.py:15-60.
    `start`: length-S start distribution.  Returns a dict of device tensors:
    states [n, max_len+1] / actions [n, max_len] (store=True), lengths [n], visit_counts [S],
    start_counts [S], n_truncated (python int; reading it synchronises)."""
    torch = require_cuda()
    S, A = tables.S, tables.A
    if tables.n_tables != 1:
        raise EngineError("sample_trajectories takes the tables of one world (use tables.select(b))")
    pol = to_device(policy)
    if tuple(pol.shape) != (S, A):
        raise EngineError("policy must have shape [S, A]")
    sd = to_device(start)
    if tuple(sd.shape) != (S,):
        raise EngineError("start must be a length-S distribution")
    cdf = torch.cumsum(sd, 0)
    mask = to_device(terminal_mask_t, torch.uint8)
    max_len = int(max_len) if max_len is not None else max(1000, 50 * S)
    dev = pol.device
    states = torch.empty((n, max_len + 1), dtype=torch.int32, device=dev) if store else None
    actions = torch.empty((n, max_len), dtype=torch.int32, device=dev) if store else None
    lengths = torch.empty(n, dtype=torch.int32, device=dev)
    visits = torch.zeros(S, dtype=torch.float64, device=dev)
    starts = torch.zeros(S, dtype=torch.float64, device=dev)
    trunc = torch.zeros(1, dtype=torch.int32, device=dev)
    ct = tables.c_struct(1)
    with _timed("sample_trajectories"):
        _check(_lib.irlb200_sample_trajectories(ctypes.byref(ct), _ptr(pol), _ptr(cdf), _ptr(mask), int(n), max_len,
                                                ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), _ptr(states), _ptr(actions),
                                                _ptr(lengths), _ptr(visits), _ptr(starts), _ptr(trunc), _stream()))
    return {"states": states, "actions": actions, "lengths": lengths, "visit_counts": visits,
            "start_counts": starts, "n_truncated": int(trunc.item()), "max_len": max_len, "start_cdf": cdf}


def features_dot(features, theta):
    """reward = features . theta (maxent.py:244), dense features [S,F] on the device."""
    torch = require_cuda()
    S, F = features.shape
    out = torch.empty(S, dtype=torch.float64, device=features.device)
    _check(_lib.irlb200_features_dot(_ptr(features), S, F, _ptr(theta), _ptr(out), _stream()))
    return out


def features_grad(features, svf_t, e_features):
    """grad = e_features - features^T . svf (maxent.py:248)."""
    torch = require_cuda()
    S, F = features.shape
    out = torch.empty(F, dtype=torch.float64, device=features.device)
    _check(_lib.irlb200_features_grad(_ptr(features), S, F, _ptr(svf_t), _ptr(e_features), _ptr(out), _stream()))
    return out


def irl_small(tables, theta, e_features, p_initial, terminal_mask_t, opt_kind, rates, eps, steps, done,
              causal=False, phi=None, discount=0.0, eps_lap=1e-5, eps_svf=1e-5, n_backward=None, max_sweeps=None):
    """The outer loop of irl / irl_causal on the device for up to `rates.shape[-1]` steps (see
    irlb200_irl_small in the header).  theta [B,S], steps [B], done [B] are updated in place;
    returns the {policy, forward} sweep counts of the last step [B,2]."""
    torch = require_cuda()
    S = tables.S
    B = int(theta.shape[0])
    ef, efshared = _maybe_shared(e_features, S, B)
    p0, p0shared = _maybe_shared(p_initial, S, B)
    mask, mshared = _maybe_shared(terminal_mask_t, S, B, torch.uint8)
    ph = None
    if causal:
        ph, pshared = _maybe_shared(phi, S, B)
        if pshared != mshared:
            raise EngineError("phi and terminal mask must both be shared or both per-problem")
    rates = to_device(rates)
    lr_shared = 1 if rates.dim() == 1 else 0
    n_rates = int(rates.shape[-1])
    counts = torch.zeros((B, 2), dtype=torch.int32, device=theta.device)
    ct = tables.c_struct(_tables_shared(tables, B))
    ms = DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
    with _timed("irl_small"):
        _check(_lib.irlb200_irl_small(
            ctypes.byref(ct), B, 1 if causal else 0, _ptr(theta), _ptr(ef), efshared, _ptr(p0), p0shared, _ptr(mask),
            _ptr(ph), mshared, 2 * S if n_backward is None else int(n_backward), float(discount), float(eps_lap),
            float(eps_svf), ms, int(opt_kind), _ptr(rates), lr_shared, n_rates, float(eps), _ptr(steps), _ptr(done),
            _ptr(counts), _stream()))
    return counts


# ---------------------------------------------------------------------------
# batched dense path: B candidates sharing one dense table (csrc/dense_batch.cu)
# ---------------------------------------------------------------------------

class DenseTables:
    """A dense p_transition[S, S', A] packed for the FP64 tensor-core contraction (irlb200_dense_pack).
    The handle owns one work buffer, reused by its calls: use a handle from one stream at a time."""

    def __init__(self, p_transition):
        torch = require_cuda()
        P = to_device(p_transition)
        if P.dim() != 3 or P.shape[0] != P.shape[1]:
            raise EngineError("p_transition must have shape [S, S, A]")
        self.S, self.A = int(P.shape[0]), int(P.shape[2])
        self.packed = torch.empty(int(_lib.irlb200_dense_pack_doubles(self.S, self.A)), dtype=torch.float64,
                                  device=P.device)
        _check(_lib.irlb200_dense_pack(_ptr(P), self.S, self.A, _ptr(self.packed), _stream()))
        self._work = None

    @property
    def shape(self):
        return (self.S, self.S, self.A)

    def work(self, B):
        torch = _torch()
        need = int(_lib.irlb200_dense_batch_work_bytes(self.S, self.A, B))
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.uint8, device=self.packed.device)
        return self._work


def dense_backward(dt, terminal_mask_t, reward, n_sweeps=None):
    """local_action_probabilities for B candidates over a shared dense table: policy [B, S, A]."""
    torch = require_cuda()
    r, B = _batch2d(reward, dt.S)
    mask = to_device(terminal_mask_t, torch.uint8)
    pol = torch.empty((B, dt.S, dt.A), dtype=torch.float64, device=r.device)
    it = torch.empty((B, dt.S), dtype=torch.float64, device=r.device)
    w = dt.work(B)
    with _timed("dense_backward"):
        _check(_lib.irlb200_dense_batch_backward(_ptr(dt.packed), dt.S, dt.A, B, _ptr(r), _ptr(mask),
                                                 2 * dt.S if n_sweeps is None else int(n_sweeps), _ptr(pol), _ptr(it),
                                                 _ptr(w), w.numel(), _stream()))
    return pol


def dense_soft_vi(dt, phi, reward, discount, eps=1e-5, max_sweeps=None):
    """local_causal_action_probabilities for B candidates: (policy [B, S, A], value [B, S])."""
    global last_info
    torch = require_cuda()
    r, B = _batch2d(reward, dt.S)
    ph = to_device(phi)
    pol = torch.empty((B, dt.S, dt.A), dtype=torch.float64, device=r.device)
    val = torch.empty((B, dt.S), dtype=torch.float64, device=r.device)
    n_iter = torch.zeros(B, dtype=torch.int32, device=r.device)
    status = torch.zeros(B, dtype=torch.int32, device=r.device)
    w = dt.work(B)
    ms = DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
    with _timed("dense_soft_vi"):
        _check(_lib.irlb200_dense_batch_succ(1, _ptr(dt.packed), dt.S, dt.A, B, _ptr(r), _ptr(ph), float(discount),
                                             float(eps), ms, 0, _ptr(val), _ptr(pol), _ptr(n_iter), _ptr(status),
                                             _ptr(w), w.numel(), _stream()))
    last_info = SweepInfo(n_iter, status)
    return pol, val


def dense_value_iteration(dt, reward, discount, eps=1e-3, max_sweeps=None, mean=False):
    global last_info
    torch = require_cuda()
    r, B = _batch2d(reward, dt.S)
    val = torch.empty((B, dt.S), dtype=torch.float64, device=r.device)
    n_iter = torch.zeros(B, dtype=torch.int32, device=r.device)
    status = torch.zeros(B, dtype=torch.int32, device=r.device)
    w = dt.work(B)
    ms = DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
    with _timed("dense_value_iteration"):
        _check(_lib.irlb200_dense_batch_succ(2, _ptr(dt.packed), dt.S, dt.A, B, _ptr(r), None, float(discount),
                                             float(eps), ms, 1 if mean else 0, _ptr(val), None, _ptr(n_iter),
                                             _ptr(status), _ptr(w), w.numel(), _stream()))
    last_info = SweepInfo(n_iter, status)
    return val


def dense_svf(dt, p_initial, terminal_mask_t, policy, eps=1e-5, max_sweeps=None, e_features=None):
    """expected_svf_from_policy for B policies [B, S, A] over a shared dense table."""
    global last_info
    torch = require_cuda()
    pol = to_device(policy)
    if pol.dim() == 2:
        pol = pol.unsqueeze(0)
    B = int(pol.shape[0])
    if tuple(pol.shape[1:]) != (dt.S, dt.A):
        raise EngineError("policy must have shape [S, A] or [B, S, A]")
    p0, p0shared = _maybe_shared(p_initial, dt.S, B)
    mask = to_device(terminal_mask_t, torch.uint8)
    out = torch.empty((B, dt.S), dtype=torch.float64, device=pol.device)
    grad, ef, efshared = None, None, 1
    if e_features is not None:
        ef, efshared = _maybe_shared(e_features, dt.S, B)
        grad = torch.empty((B, dt.S), dtype=torch.float64, device=pol.device)
    n_iter = torch.zeros(B, dtype=torch.int32, device=pol.device)
    status = torch.zeros(B, dtype=torch.int32, device=pol.device)
    w = dt.work(B)
    ms = DEFAULT_MAX_SWEEPS if max_sweeps is None else int(max_sweeps)
    with _timed("dense_svf"):
        _check(_lib.irlb200_dense_batch_svf(_ptr(dt.packed), dt.S, dt.A, B, _ptr(p0), p0shared, _ptr(mask), _ptr(pol),
                                            float(eps), ms, _ptr(out), _ptr(ef), efshared, _ptr(grad), _ptr(n_iter),
                                            _ptr(status), _ptr(w), w.numel(), _stream()))
    last_info = SweepInfo(n_iter, status)
    return (out, grad) if grad is not None else out
