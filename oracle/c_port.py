"""ctypes binding of oracle/c/irl_oracle.c (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

The C restatement works on per-(state, action) successor lists; `ell_from_dense` / `ell_from_sparse`
build them from the reference's dense table or from the scipy restatement's CSR matrices.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libirl_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = ctypes.CDLL(_LIB)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def ell_from_dense(P):
    """(sidx [S,A,K] int32, sp [S,A,K] f64) with successors ascending, unused slots p = 0, idx = s."""
    S, _, A = P.shape
    K = max(1, int((P != 0).sum(axis=1).max()))
    sidx = np.repeat(np.arange(S, dtype=np.int32)[:, None, None], A, axis=1).repeat(K, axis=2).copy()
    sp = np.zeros((S, A, K))
    for s in range(S):
        for a in range(A):
            nz = np.nonzero(P[s, :, a])[0]
            sidx[s, a, :len(nz)] = nz
            sp[s, a, :len(nz)] = P[s, nz, a]
    return sidx, sp


def ell_from_sparse(mdp):
    """Same from a sparse_port.SparseMDP (per-action CSR)."""
    S, A = mdp.n_states, mdp.n_actions
    K = max(int(np.diff(m.indptr).max()) for m in mdp.per_action)
    sidx = np.repeat(np.arange(S, dtype=np.int32)[:, None, None], A, axis=1).repeat(K, axis=2).copy()
    sp = np.zeros((S, A, K))
    for a, m in enumerate(mdp.per_action):
        m = m.tocsr()
        m.sort_indices()
        cnt = np.diff(m.indptr)
        rows = np.repeat(np.arange(S), cnt)
        slot = np.arange(m.nnz) - np.repeat(m.indptr[:-1], cnt)
        sidx[rows, a, slot] = m.indices
        sp[rows, a, slot] = m.data
    return sidx, sp


def _mask(terminal, S):
    m = np.zeros(S, dtype=np.uint8)
    m[np.asarray(list(terminal), dtype=np.int64)] = 1
    return m


def backward(sidx, sp, terminal, reward, n_sweeps=None, rescale=False):
    S, A, K = sp.shape
    pol = np.empty((S, A))
    reward = np.ascontiguousarray(reward, dtype=np.float64)
    rc = load().oracle_backward(S, A, K, _p(sidx, ctypes.c_int), _p(sp, ctypes.c_double),
                                _p(_mask(terminal, S), ctypes.c_ubyte), _p(reward, ctypes.c_double),
                                ctypes.c_long(2 * S if n_sweeps is None else n_sweeps), int(rescale),
                                _p(pol, ctypes.c_double))
    assert rc == 0
    return pol


def soft_vi(sidx, sp, phi, reward, discount, eps=1e-5, max_sweeps=0):
    S, A, K = sp.shape
    pol, val, n = np.empty((S, A)), np.empty(S), ctypes.c_long(0)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    reward = np.ascontiguousarray(reward, dtype=np.float64)
    rc = load().oracle_soft_vi(S, A, K, _p(sidx, ctypes.c_int), _p(sp, ctypes.c_double), _p(phi, ctypes.c_double),
                               _p(reward, ctypes.c_double), ctypes.c_double(discount), ctypes.c_double(eps),
                               ctypes.c_long(max_sweeps), _p(pol, ctypes.c_double), _p(val, ctypes.c_double),
                               ctypes.byref(n))
    assert rc == 0
    return pol, val, n.value


def svf(sidx, sp, p0, terminal, policy, eps=1e-5, max_sweeps=0, threads=None):
    """threads: None = the OpenMP gather form for S >= 4096 (bitwise the serial loop, see irl_oracle.c),
    True / False force it."""
    S, A, K = sp.shape
    d, n = np.empty(S), ctypes.c_long(0)
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    policy = np.ascontiguousarray(policy, dtype=np.float64)
    mt = (S >= 4096) if threads is None else bool(threads)
    fn = load().oracle_svf_mt if mt else load().oracle_svf
    rc = fn(S, A, K, _p(sidx, ctypes.c_int), _p(sp, ctypes.c_double), _p(p0, ctypes.c_double),
                           _p(_mask(terminal, S), ctypes.c_ubyte), _p(policy, ctypes.c_double),
                           ctypes.c_double(eps), ctypes.c_long(max_sweeps), _p(d, ctypes.c_double), ctypes.byref(n))
    assert rc == 0
    return d, n.value


def value_iteration(sidx, sp, reward, discount, eps=1e-3, max_sweeps=0):
    S, A, K = sp.shape
    v, n = np.empty(S), ctypes.c_long(0)
    reward = np.ascontiguousarray(reward, dtype=np.float64)
    rc = load().oracle_vi(S, A, K, _p(sidx, ctypes.c_int), _p(sp, ctypes.c_double), _p(reward, ctypes.c_double),
                          ctypes.c_double(discount), ctypes.c_double(eps), ctypes.c_long(max_sweeps),
                          _p(v, ctypes.c_double), ctypes.byref(n))
    assert rc == 0
    return v, n.value


def batch_maxent_step(sidx_b, sp_b, terminal, p0, rewards, eps=1e-5, max_sweeps=0):
    """B worlds ([B,S,A,K] tables, [B,S] rewards) on all OpenMP threads: (svf [B,S], forward sweeps [B])."""
    B, S, A, K = sp_b.shape
    out = np.empty((B, S))
    n = np.zeros(B, dtype=np.int64)
    sidx_b = np.ascontiguousarray(sidx_b, dtype=np.int32)
    sp_b = np.ascontiguousarray(sp_b, dtype=np.float64)
    rewards = np.ascontiguousarray(rewards, dtype=np.float64)
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    rc = load().oracle_batch_maxent_step(B, S, A, K, _p(sidx_b, ctypes.c_int), _p(sp_b, ctypes.c_double),
                                         _p(_mask(terminal, S), ctypes.c_ubyte), _p(p0, ctypes.c_double),
                                         _p(rewards, ctypes.c_double), ctypes.c_double(eps), ctypes.c_long(max_sweeps),
                                         _p(out, ctypes.c_double), _p(n, ctypes.c_long))
    assert rc == 0
    return out, n
