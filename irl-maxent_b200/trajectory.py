"""
trajectory -- expert demonstrations with the interface of the reference's
`trajectory.py` (`/root/reference/src/trajectory.py`).  The reference's functions
stay host side and draw from numpy's global generator in the same order as the
reference, so a seeded run produces the same trajectories.
`generate_trajectories_device` is the device path for worlds where a Python loop
per step is too slow: one CUDA thread per trajectory,
counter-based generator, visit statistics accumulated on the fly.
"""

import numpy as np


class Trajectory:
    """A list of (state_from, action, state_to) transitions (reference: trajectory.py:10-49)."""

    def __init__(self, transitions):
        self._t = transitions

    def transitions(self):
        return self._t

    def states(self):
        """Visited states in order, the final `state_to` included."""
        return iter([t[0] for t in self._t] + [self._t[-1][2]])

    def __repr__(self):
        return "Trajectory({})".format(repr(self._t))

    def __str__(self):
        return "{}".format(self._t)


def _choice_sparse(states, probs):
    """`np.random.choice(range(S), p=row)` for a row given by its non-zero entries (ascending).
    numpy draws ONE uniform and searches the normalised cumulative sum from the right; zero entries
    only repeat cdf values, so searching the cdf of the non-zeros picks the same state and leaves
    the global generator in the same state as the dense call -- seeded runs stay bit-identical."""
    cdf = np.cumsum(probs)
    cdf /= cdf[-1]
    u = np.random.random_sample()
    return int(states[np.searchsorted(cdf, u, side='right')])


def generate_trajectory(world, policy, start, final):
    """Roll `policy` out from `start` until a state in `final` (reference: trajectory.py:52-87).
    Worlds that expose `successors(s, a)` are sampled from their sparse rows in O(1) per step
    instead of an O(S) dense row (the dense table of a large world does not exist)."""
    state, steps = start, []
    sparse = hasattr(world, "successors")
    states = range(world.n_states)
    while state not in final:
        action = policy(state)
        if sparse:
            nxt = _choice_sparse(*world.successors(state, action))
        else:
            nxt = np.random.choice(states, p=world.p_transition[state, :, action])
        steps.append((state, action, nxt))
        state = nxt
    return Trajectory(steps)


def generate_trajectories(n, world, policy, start, final):
    """Generator of n trajectories; `start` is a state, a list of states, or a
    length-S start distribution (reference: trajectory.py:90-128)."""
    starts = np.atleast_1d(start)

    def one():
        if len(starts) == world.n_states:
            s = np.random.choice(range(world.n_states), p=starts)
        else:
            s = np.random.choice(starts)
        return generate_trajectory(world, policy, s, final)

    return (one() for _ in range(n))


def policy_adapter(policy):
    """Deterministic policy array -> callable (reference: trajectory.py:131-147)."""
    return lambda state: policy[state]


def stochastic_policy_adapter(policy):
    """Stochastic policy [S, A] -> sampling callable (reference: trajectory.py:150-169)."""
    return lambda state: np.random.choice([*range(policy.shape[1])], p=policy[state, :])


class DeviceTrajectories:
    """n rollouts held on the device (`_irlb200.sample_trajectories`).

    Iterating yields host `Trajectory` objects (so everything written against the reference's
    interface keeps working); `maxent.feature_expectation_from_trajectories` and
    `maxent.initial_probabilities_from_trajectories` take the device visit / start counts
    directly instead of looping over states in Python."""

    def __init__(self, raw, n_states):
        self.n_states = n_states
        self.states, self.actions, self.lengths = raw["states"], raw["actions"], raw["lengths"]
        self.visit_counts, self.start_counts = raw["visit_counts"], raw["start_counts"]
        self.n_truncated, self.max_len = raw["n_truncated"], raw["max_len"]

    def __len__(self):
        return int(self.lengths.numel())

    def __iter__(self):
        if self.states is None:
            raise ValueError("trajectories were sampled with store=False: only the statistics exist")
        st, ac, ln = self.states.cpu().numpy(), self.actions.cpu().numpy(), self.lengths.cpu().numpy()
        for i in range(len(ln)):
            k = int(ln[i])
            yield Trajectory([(int(st[i, j]), int(ac[i, j]), int(st[i, j + 1])) for j in range(k)])


def generate_trajectories_device(n, world, policy, start, final, seed=0, max_len=None, store=True):
    """`generate_trajectories` (reference: trajectory.py:90-128) on the device.

    `world`: a world with `.tables()` / a `_irlb200.Tables` handle / a dense `p_transition`;
    `policy`: stochastic policy array [S, A] (the array the reference wraps with
    `stochastic_policy_adapter`) or a deterministic policy [S] of action indices
    (`policy_adapter`); `start`: a state, a list of states (uniform) or a length-S start
    distribution, as in the reference; `final`: terminal states.  The random stream is the
    device generator's (see include/irl_maxent_b200.h), not numpy's."""
    import _irlb200 as E
    tables = world.tables() if hasattr(world, "tables") else E.as_tables(world)
    S, A = tables.S, tables.A
    if E.is_tensor(policy):
        pol = policy
        if pol.dim() == 1:
            pol = E._torch().nn.functional.one_hot(pol.long(), A).double()
    else:
        pol = np.asarray(policy)
        if pol.ndim == 1:
            pol = np.eye(A)[pol.astype(np.int64)]
    starts = np.atleast_1d(start)
    if len(starts) == S:
        dist = starts.astype(float)
    else:
        dist = np.zeros(S)
        np.add.at(dist, starts.astype(np.int64), 1.0 / len(starts))
    raw = E.sample_trajectories(tables, pol, dist, E.terminal_mask(final, S), n, seed, max_len=max_len, store=store)
    return DeviceTrajectories(raw, S)
