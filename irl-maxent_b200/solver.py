"""
solver -- MDP solvers with the interface of the reference's `solver.py`
(`/root/reference/src/solver.py`).  The Bellman sweeps run in the sm_100a
kernels (`irlb200_value_iteration`); the policy-extraction helpers are
vectorised gathers over the world's successor structure instead of the
reference's Python double loops.
"""

import numpy as np

import _irlb200 as E


def _out(t, like):
    return t if E.is_tensor(like) else t.cpu().numpy()


def value_iteration(p, reward, discount, eps=1e-3):
    """v = reward + discount * max_a P_a v until max|dv| <= eps (reference: solver.py:9-52).
    Returns an array of shape (S,)."""
    tables = E.as_tables(p)
    v = E.value_iteration(tables, reward, discount, eps)
    return _out(v[0], reward)


def stochastic_value_iteration(p, reward, discount, eps=1e-3):
    """As value_iteration with the mean over actions instead of the max
    (reference: solver.py:55-104)."""
    tables = E.as_tables(p)
    v = E.value_iteration(tables, reward, discount, eps, mean=True)
    return _out(v[0], reward)


def _intended_next_states(world):
    """[S, A] table of world.state_index_transition(s, a) (reference: gridworld.py:106-122)."""
    if hasattr(world, "intended_next_states"):
        return world.intended_next_states()
    return np.array([[world.state_index_transition(s, a) for a in range(world.n_actions)]
                     for s in range(world.n_states)])


def optimal_policy_from_value(world, value):
    """Greedy action w.r.t. the value of the intended successor (reference: solver.py:107-126)."""
    v = value.cpu().numpy() if E.is_tensor(value) else np.asarray(value)
    return np.argmax(v[_intended_next_states(world)], axis=1)


def optimal_policy(world, reward, discount, eps=1e-3):
    """value_iteration followed by optimal_policy_from_value (reference: solver.py:129-152)."""
    return optimal_policy_from_value(world, value_iteration(world.p_transition, reward, discount, eps))


def stochastic_policy_from_value(world, value, w=lambda x: x):
    """p(a|s) proportional to w(value of the intended successor) (reference: solver.py:155-181)."""
    v = value.cpu().numpy() if E.is_tensor(value) else np.asarray(value)
    nxt = _intended_next_states(world)
    weights = np.array([[w(v[nxt[s, a]]) for a in range(nxt.shape[1])] for s in range(nxt.shape[0])])
    return weights / np.sum(weights, axis=1)[:, None]
