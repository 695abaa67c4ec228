"""
Device trajectory sampler (csrc/trajectories.cu) against its CPU
restatement (oracle/sampler_port.py): trajectories bit-identical (index work), statistics exact,
and the reference-facing wrappers (`trajectory.generate_trajectories_device`, the trajectory
statistics of maxent.py:15-60) consistent with the host path.  Needs a B200: `pytest -m gpu`.
"""
import numpy as np
import pytest

import _irlb200 as E
import gridworld as W
import maxent as M
import optimizer as O
import solver as S
import trajectory as T

from oracle import sampler_port as SPL

pytestmark = pytest.mark.gpu


def _ell_rows(tables):
    idx = tables.succ_idx[0].cpu().numpy()          # [Ks][S]
    p = tables.succ_p[0].cpu().numpy()              # [A][Ks][S]
    return lambda s, a: (idx[:, s], p[a, :, s])


def _random_policy(rng, S, A):
    pol = rng.random((S, A)) + 0.05
    pol[rng.random((S, A)) < 0.2] = 0.0             # some zero-probability actions
    pol[pol.sum(axis=1) == 0.0, 0] = 1.0
    return pol / pol.sum(axis=1, keepdims=True)


@pytest.mark.parametrize("n,p_slip,seed", [(5, 0.2, 1), (8, 0.35, 2 ** 40 + 17), (12, 0.05, 3)])
def test_device_trajectories_equal_the_oracle_bit_for_bit(n, p_slip, seed):
    S_ = n * n
    rng = np.random.default_rng(n)
    tables = E.gridworld_tables(n, p_slip)
    pol = _random_policy(rng, S_, 4)
    start = rng.random(S_); start[rng.random(S_) < 0.5] = 0.0; start[0] += 0.1; start /= start.sum()
    final = [S_ - 1, S_ // 2]
    raw = E.sample_trajectories(tables, pol, start, E.terminal_mask(final, S_), 96, seed, max_len=400)
    st, ac, ln = raw["states"].cpu().numpy(), raw["actions"].cpu().numpy(), raw["lengths"].cpu().numpy()
    cdf = raw["start_cdf"].cpu().numpy()
    succ = _ell_rows(tables)
    n_trunc = 0
    visits, starts = np.zeros(S_), np.zeros(S_)
    for i in range(96):
        o_st, o_ac, trunc = SPL.sample_trajectory(succ, pol, cdf, final, i, seed, 400)
        n_trunc += trunc
        assert ln[i] == len(o_ac)
        assert st[i, :ln[i] + 1].tolist() == o_st and ac[i, :ln[i]].tolist() == o_ac
        np.add.at(visits, o_st, 1.0)
        starts[o_st[0]] += 1.0
    assert raw["n_truncated"] == n_trunc
    assert np.array_equal(raw["visit_counts"].cpu().numpy(), visits)
    assert np.array_equal(raw["start_counts"].cpu().numpy(), starts)


def test_random_non_grid_mdp_and_deterministic_policy(golden):
    P = golden("random_mdps")["r2_P"]
    S_, _, A = P.shape
    tables = E.compress_dense(P)
    rng = np.random.default_rng(9)
    det = rng.integers(0, A, S_)
    dt = T.generate_trajectories_device(40, tables, det, [0, 1, 2], [S_ - 1], seed=5, max_len=300)
    succ = _ell_rows(tables)
    onehot = np.eye(A)[det]
    st, ln = dt.states.cpu().numpy(), dt.lengths.cpu().numpy()
    start = np.zeros(S_); start[[0, 1, 2]] = 1.0 / 3
    cdf = np.cumsum(start)
    for i in range(40):
        o_st, o_ac, _ = SPL.sample_trajectory(succ, onehot, cdf, [S_ - 1], i, 5, 300)
        assert st[i, :ln[i] + 1].tolist() == o_st
        assert all(a == det[s] for s, a in zip(o_st[:-1], o_ac))
        for (s, a, s2) in zip(o_st[:-1], o_ac, o_st[1:]):
            assert P[s, s2, a] > 0.0                     # only possible transitions are taken


def test_statistics_match_the_host_functions_on_the_same_trajectories():
    world = W.IcyGridWorld(size=5, p_slip=0.2)
    reward = np.zeros(25); reward[24] = 1.0; reward[8] = 0.65
    value = S.value_iteration(world.p_transition, reward, 0.7)
    policy = S.stochastic_policy_from_value(world, value, w=lambda x: x ** 5)
    dt = T.generate_trajectories_device(200, world, policy, 0, [24], seed=7)
    assert dt.n_truncated == 0 and len(dt) == 200
    host = list(dt)                                      # reference-style Trajectory objects
    assert all(t.transitions()[0][0] == 0 and t.transitions()[-1][2] == 24 for t in host)
    ident = W.state_features(world)
    assert np.array_equal(M.feature_expectation_from_trajectories(ident, dt),
                          M.feature_expectation_from_trajectories(ident, host))
    assert np.array_equal(M.initial_probabilities_from_trajectories(25, dt),
                          M.initial_probabilities_from_trajectories(25, host))
    coord = W.coordinate_features(world)
    np.testing.assert_allclose(M.feature_expectation_from_trajectories(coord, dt),
                               M.feature_expectation_from_trajectories(coord, host), rtol=1e-13)
    # the whole IRL run is the same whichever container carries the demonstrations
    def run(tj):
        return M.irl(world.p_transition, ident, [24], tj, O.ExpSga(lr=O.linear_decay(lr0=0.2)), O.Constant(1.0))
    assert np.array_equal(run(dt), run(host))


def test_transition_frequencies_follow_the_table():
    """First step out of state 12 of the 5x5 world under a fixed policy row: empirical (action,
    successor) frequencies of 20 000 rollouts vs policy * P, 5-sigma bound per cell."""
    world = W.IcyGridWorld(size=5, p_slip=0.3)
    P = world.p_transition
    pol = np.tile(np.array([0.1, 0.2, 0.3, 0.4]), (25, 1))
    n = 20000
    dt = T.generate_trajectories_device(n, world, pol, 12, [24], seed=11, max_len=1)
    st, ac = dt.states.cpu().numpy(), dt.actions.cpu().numpy()
    assert dt.n_truncated == n and (st[:, 0] == 12).all()
    for a in range(4):
        for s2 in range(25):
            expect = pol[12, a] * P[12, s2, a]
            got = np.mean((ac[:, 0] == a) & (st[:, 1] == s2))
            assert abs(got - expect) <= 5 * np.sqrt(max(expect * (1 - expect), 1e-12) / n) + 1e-12


def test_large_world_statistics_only():
    """128 x 128: 256 goal-directed rollouts without storing them (the dense row the reference
    samples from would need the 8.6 GB table)."""
    n = 128; S_ = n * n
    world = W.IcyGridWorld(size=n, p_slip=0.2)          # lazy: no dense table above 4 096 states
    r = np.full(S_, -0.1); r[S_ - 1] = 1.0
    tables = world.tables()
    pol = E.soft_vi(tables, E.terminal_phi([S_ - 1], S_), r, 0.9)[0]
    dt = T.generate_trajectories_device(256, world, pol, 0, [S_ - 1], seed=3, store=False)
    assert dt.n_truncated == 0
    fe = M.feature_expectation_from_trajectories(W.state_features(world), dt)
    p0 = M.initial_probabilities_from_trajectories(S_, dt)
    assert p0[0] == 1.0 and p0.sum() == 1.0
    assert fe[S_ - 1] == 1.0 and fe.sum() >= 2 * (n - 1) + 1          # every rollout ends in the goal
    with pytest.raises(ValueError):
        list(dt)
