"""
trajectory -- expert demonstrations with the interface of the reference's
`trajectory.py` (`/root/reference/src/trajectory.py`).  Host side; draws from
numpy's global generator in the same order as the reference, so a seeded run
produces the same trajectories.
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
