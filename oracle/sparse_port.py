"""
Sparse (scipy CSR) restatement of the reference hot path for state counts the
dense `[S,S',A]` table cannot reach (128x128: 8.6 GB, 2048x2048: 141 TB).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Same loops as oracle/dense_port.py (and therefore as maxent.py:63-114,
:119-159, :279-341 and solver.py:9-52) with every `P_a.dot(x)` replaced by a
CSR mat-vec over the non-zeros of `P_a`.  Not the reference's cost profile --
bench.py reports it as the "fair sparse CPU" line, never as the reference.

Parity: pinned indirectly -- tests/test_oracle_golden.py checks it against the
reference-generated fixtures at the sizes where the dense table exists
(differences are last-ulp summation-order effects, bound 1e-12 relative) and
checks iteration counts exactly.
"""

import numpy as np
import scipy.sparse as sp

from .dense_port import ACTIONS, softmax, terminal_reward, _NEG_HUGE


class SparseMDP:
    """Per-action CSR matrices P_a[s, s'] plus their transposes."""

    def __init__(self, per_action):
        self.per_action = [m.tocsr() for m in per_action]
        self.n_actions = len(per_action)
        self.n_states = per_action[0].shape[0]
        self._T = None

    @classmethod
    def from_dense(cls, p_transition):
        return cls([sp.csr_matrix(p_transition[:, :, a]) for a in range(p_transition.shape[2])])

    def transposed(self):
        if self._T is None:
            self._T = [m.T.tocsr() for m in self.per_action]
        return self._T


def icy_gridworld_sparse(size, p_slip=0.2):
    """IcyGridWorld(size, p_slip) as per-action CSR without the dense detour.
    Value cases as in gridworld.py:200-248 (see dense_port.icy_gridworld_table)."""
    n, A = size, len(ACTIONS)
    S = n * n
    p = float(p_slip)
    v_intended = 1.0 - p + p / A
    v_slip = p / A
    v_stay_corner_into = 1.0 - p + 2.0 * p / A
    v_stay_edge_into = 1.0 - p + p / A
    v_stay_corner = 2.0 * p / A
    v_stay_edge = p / A

    s = np.arange(S)
    fx, fy = s % n, s // n
    xb = (fx == 0) | (fx == n - 1)
    yb = (fy == 0) | (fy == n - 1)
    corner, edge = xb & yb, xb | yb
    mats = []
    for a, (ax, ay) in enumerate(ACTIONS):
        rows, cols, vals = [], [], []
        for b, (dx, dy) in enumerate(ACTIONS):
            tx, ty = fx + dx, fy + dy
            ok = (tx >= 0) & (tx < n) & (ty >= 0) & (ty < n)
            rows.append(s[ok])
            cols.append((ty * n + tx)[ok])
            vals.append(np.full(ok.sum(), v_intended if b == a else v_slip))
        tx, ty = fx + ax, fy + ay
        into_wall = ~((tx >= 0) & (tx < n) & (ty >= 0) & (ty < n))
        stay = np.zeros(S)
        stay[into_wall & corner] = v_stay_corner_into
        stay[into_wall & ~corner] = v_stay_edge_into
        stay[~into_wall & corner] = v_stay_corner
        stay[~into_wall & ~corner & edge] = v_stay_edge
        nz = stay != 0.0
        rows.append(s[nz]); cols.append(s[nz]); vals.append(stay[nz])
        m = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(S, S))
        m.sort_indices()
        mats.append(m)
    return SparseMDP(mats)


def expected_svf_from_policy(mdp, p_initial, terminal, p_action, eps=1e-5, max_sweeps=None):
    """maxent.py:63-114 over CSR.  Returns (d, n_sweeps)."""
    S, A = mdp.n_states, mdp.n_actions
    keep = np.ones(S)
    keep[terminal] = 0.0                        # :99 zero the outgoing rows of terminal states
    PT = mdp.transposed()
    d = np.zeros(S)
    delta, n = np.inf, 0
    while delta > eps:
        parts = [PT[a].dot(keep * (p_action[:, a] * d)) for a in range(A)]      # :109
        d_new = p_initial + np.array(parts).sum(axis=0)                          # :110
        delta, d = np.max(np.abs(d_new - d)), d_new                              # :112
        n += 1
        if max_sweeps is not None and n >= max_sweeps:
            break
    return d, n


def local_action_probabilities(mdp, terminal, reward, rescale=True):
    """maxent.py:119-159 over CSR, range-extended by default (see dense_port)."""
    S, A = mdp.n_states, mdp.n_actions
    er = np.exp(reward)
    zs = np.zeros(S)
    zs[terminal] = 1.0
    za = None
    for _ in range(2 * S):
        za = np.array([er * mdp.per_action[a].dot(zs) for a in range(A)]).T
        zs = za.sum(axis=1)
        if rescale:
            m = np.max(zs)
            if np.isfinite(m) and m > 0.0:
                _, e = np.frexp(m)
                zs = np.ldexp(zs, -int(e))
                za = np.ldexp(za, -int(e))
    return za / zs[:, None]


def local_causal_action_probabilities(mdp, terminal, reward, discount, eps=1e-5, max_sweeps=None,
                                      return_value=False):
    """maxent.py:279-341 over CSR.  Returns (policy, n_sweeps) [, v with return_value]."""
    S, A = mdp.n_states, mdp.n_actions
    phi = terminal_reward(terminal, S)
    v = _NEG_HUGE * np.ones(S)
    delta, n, q = np.inf, 0, None
    with np.errstate(over='ignore', invalid='ignore'):
        while delta > eps:
            v_old = v
            q = np.array([reward + discount * mdp.per_action[a].dot(v_old) for a in range(A)]).T
            v = phi
            for a in range(A):
                v = softmax(v, q[:, a])
            v = np.array(v, dtype=float)
            delta = np.max(np.abs(v - v_old))
            n += 1
            if max_sweeps is not None and n >= max_sweeps:
                break
        if return_value:
            return np.exp(q - v[:, None]), n, v
        return np.exp(q - v[:, None]), n


def value_iteration(mdp, reward, discount, eps=1e-3, max_sweeps=None):
    """solver.py:9-52 over CSR.  Returns (v, n_sweeps)."""
    S, A = mdp.n_states, mdp.n_actions
    v = np.zeros(S)
    delta, n = np.inf, 0
    while delta > eps:
        v_old = v
        q = discount * np.array([mdp.per_action[a].dot(v) for a in range(A)])
        v = reward + np.max(q, axis=0)
        delta = np.max(np.abs(v_old - v))
        n += 1
        if max_sweeps is not None and n >= max_sweeps:
            break
    return v, n
