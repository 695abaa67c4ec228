"""
gridworld -- GridWorld / IcyGridWorld MDPs with the interface of the reference's
`gridworld.py` (`/root/reference/src/gridworld.py`).

The dense `p_transition[S, S', A]` attribute is kept (built lazily by a
loop-free numpy pass that is bit-identical to the reference's O(S^2 A) Python
loop), and `tables()` builds the compressed successor / predecessor tables
directly on the device, which is the only representation that exists for
large grids (128x128 dense is 8.6 GB, 2048x2048 would be 141 TB).
"""

import numpy as np

ACTIONS = [(1, 0), (-1, 0), (0, 1), (0, -1)]


class GridWorld:
    """Deterministic size x size grid; moving into a wall keeps the agent in place.
    State index = y * size + x (reference: gridworld.py:23-174)."""

    icy = False

    def __init__(self, size, dense=None):
        self.size = size
        self.actions = list(ACTIONS)
        self.n_states = size ** 2
        self.n_actions = len(self.actions)
        self._dense = None
        self._tables = None
        # the reference builds the dense table eagerly; do so while it is small
        if dense is True or (dense is None and self.n_states <= 4096):
            self._dense = self._transition_prob_table()

    # -- index helpers (reference: gridworld.py:54-122) --------------------------------
    def state_index_to_point(self, state):
        return state % self.size, state // self.size

    def state_point_to_index(self, state):
        return state[1] * self.size + state[0]

    def state_point_to_index_clipped(self, state):
        hi = self.size - 1
        return self.state_point_to_index((min(max(state[0], 0), hi), min(max(state[1], 0), hi)))

    def state_index_transition(self, s, a):
        x, y = self.state_index_to_point(s)
        dx, dy = self.actions[a]
        return self.state_point_to_index_clipped((x + dx, y + dy))

    def intended_next_states(self):
        """[S, A] array of state_index_transition for all (s, a), vectorised."""
        s = np.arange(self.n_states)
        x, y = s % self.size, s // self.size
        cols = []
        for dx, dy in self.actions:
            cols.append(np.clip(y + dy, 0, self.size - 1) * self.size + np.clip(x + dx, 0, self.size - 1))
        return np.stack(cols, axis=1)

    # -- transition model ----------------------------------------------------------------
    @property
    def p_transition(self):
        if self._dense is None:
            self._dense = self._transition_prob_table()
        return self._dense

    def _self_loop_values(self, a, into_wall, corner, edge):
        stay = np.zeros(self.n_states)
        stay[into_wall] = 1.0
        return stay

    def _neighbour_values(self, a):
        """(intended value, value of every other in-grid neighbour)"""
        return 1.0, 0.0

    def _transition_prob_table(self):
        """Dense [from, to, action] table (reference: gridworld.py:124-171, :200-248)."""
        n, S, A = self.size, self.n_states, self.n_actions
        table = np.zeros((S, S, A))
        s = np.arange(S)
        x, y = s % n, s // n
        xb, yb = (x == 0) | (x == n - 1), (y == 0) | (y == n - 1)
        corner, edge = xb & yb, xb | yb
        for a, (ax, ay) in enumerate(self.actions):
            v_int, v_other = self._neighbour_values(a)
            if v_other != 0.0:
                for dx, dy in self.actions:
                    tx, ty = x + dx, y + dy
                    ok = (tx >= 0) & (tx < n) & (ty >= 0) & (ty < n)
                    table[s[ok], (ty * n + tx)[ok], a] = v_other
            tx, ty = x + ax, y + ay
            inside = (tx >= 0) & (tx < n) & (ty >= 0) & (ty < n)
            table[s[inside], (ty * n + tx)[inside], a] = v_int
            table[s, s, a] = self._self_loop_values(a, ~inside, corner, edge)
        return table

    def _transition_prob(self, s_from, s_to, a):
        return self.p_transition[s_from, s_to, a]

    def successors(self, s, a):
        """Sparse row of the transition model: (states, probabilities) of the non-zero entries of
        p_transition[s, :, a], states ascending -- without touching the dense table.  The values
        are formed by the same expressions as the table, so they are bit-identical to it."""
        n = self.size
        x, y = s % n, s // n
        ax, ay = self.actions[a]
        xb, yb = x in (0, n - 1), y in (0, n - 1)
        v_int, v_other = self._neighbour_values(a)
        inside = 0 <= x + ax < n and 0 <= y + ay < n
        stay = self._self_loop_values_scalar(not inside, xb and yb, xb or yb)
        out_s, out_p = [], []
        for t, (dx, dy) in ((s - n, (0, -1)), (s - 1, (-1, 0)), (s, (0, 0)), (s + 1, (1, 0)), (s + n, (0, 1))):
            if (dx, dy) == (0, 0):
                p = stay
            elif not (0 <= x + dx < n and 0 <= y + dy < n):
                continue
            else:
                p = v_int if (dx, dy) == (ax, ay) else v_other
            if p != 0.0:
                out_s.append(t)
                out_p.append(p)
        return np.array(out_s, dtype=np.int64), np.array(out_p)

    def _self_loop_values_scalar(self, into_wall, corner, edge):
        return 1.0 if into_wall else 0.0

    def tables(self):
        """Device-resident compressed tables, built without the dense detour.  Worlds beyond the
        register-resident kernels (side > 128) are streamed from HBM every sweep and get the compact
        4-slot form; results are bitwise the same."""
        if self._tables is None:
            import _irlb200 as E
            self._tables = E.gridworld_tables(self.size, getattr(self, "p_slip", None), icy=self.icy,
                                              slots=4 if self.size > 128 else 5)
        return self._tables

    def __repr__(self):
        return "GridWorld(size={})".format(self.size)


class IcyGridWorld(GridWorld):
    """Grid world on ice: with probability p_slip the agent ends up in a random
    neighbouring cell (or stays, at walls) instead of the intended one
    (reference: gridworld.py:177-251)."""

    icy = True

    def __init__(self, size, p_slip=0.2, dense=None):
        self.p_slip = p_slip
        super().__init__(size, dense)

    def _neighbour_values(self, a):
        p, nA = self.p_slip, self.n_actions
        return 1.0 - p + p / nA, p / nA

    def _self_loop_values(self, a, into_wall, corner, edge):
        p, nA = self.p_slip, self.n_actions
        stay = np.zeros(self.n_states)
        stay[into_wall & corner] = 1.0 - p + 2.0 * p / nA
        stay[into_wall & ~corner] = 1.0 - p + p / nA
        stay[~into_wall & corner] = 2.0 * p / nA
        stay[~into_wall & ~corner & edge] = p / nA
        return stay

    def _self_loop_values_scalar(self, into_wall, corner, edge):
        p, nA = self.p_slip, self.n_actions
        if into_wall:
            return 1.0 - p + 2.0 * p / nA if corner else 1.0 - p + p / nA
        if corner:
            return 2.0 * p / nA
        return p / nA if edge else 0.0

    def __repr__(self):
        return "IcyGridWorld(size={}, p_slip={})".format(self.size, self.p_slip)


class IdentityFeatures:
    """The S x S identity feature matrix of `state_features` without the S^2 storage (2.1 GB at
    128 x 128).  `maxent.irl` / `irl_causal` recognise it and skip both feature products
    (features.dot(theta) == theta, features.T.dot(svf) == svf, bit for bit)."""

    def __init__(self, n_states):
        self.shape = (n_states, n_states)
        self.ndim = 2

    def dot(self, theta):
        return theta

    def __array__(self, dtype=None, copy=None):
        return np.identity(self.shape[0], dtype=dtype or float)


def state_features(world, implicit=None):
    """One indicator feature per state: the S x S identity (reference: gridworld.py:254-268).
    For worlds above 4 096 states (or implicit=True) an `IdentityFeatures` stand-in is returned."""
    if implicit or (implicit is None and world.n_states > 4096):
        return IdentityFeatures(world.n_states)
    return np.identity(world.n_states)


def coordinate_features(world):
    """S x size matrix with a count at the x column and at the y column of each state
    (reference: gridworld.py:271-293)."""
    features = np.zeros((world.n_states, world.size))
    s = np.arange(world.n_states)
    np.add.at(features, (s, s % world.size), 1)
    np.add.at(features, (s, s // world.size), 1)
    return features
