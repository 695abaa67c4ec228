"""The oracle (oracle/*.py) against every fixture generated from the unmodified
reference (tests/golden/generate_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import dense_port as D
from oracle import sparse_port as SP


class _Traj:
    """Duck-typed trajectory (reference: trajectory.py:10-49)."""

    def __init__(self, t):
        self._t = [tuple(int(v) for v in row) for row in t]

    def transitions(self):
        return self._t

    def states(self):
        return [x[0] for x in self._t] + [self._t[-1][2]]


def load_trajectories(g):
    out, off = [], 0
    for n in g["traj_len"]:
        out.append(_Traj(g["traj_flat"][off:off + n]))
        off += n
    return out


# ---------------------------------------------------------------- worlds ----

@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 8])
def test_world_tables_bit_identical(golden, n):
    g = golden("worlds")
    assert np.array_equal(D.gridworld_table(n), g["grid_%d" % n])
    for p in (0.2, 0.35):
        ref = g["icy_%d_%s" % (n, p)]
        assert np.array_equal(D.icy_gridworld_table(n, p), ref)
        sp_tab = SP.icy_gridworld_sparse(n, p)
        for a in range(4):
            assert np.array_equal(sp_tab.per_action[a].toarray(), ref[:, :, a])


def test_world_rows_sum_to_one(golden):
    g = golden("worlds")
    for k in g.files:
        if k.startswith("icy_1_"):
            continue        # reference quirk: the 1x1 icy world's rows sum to 1 - p_slip/2
        assert np.allclose(g[k].sum(axis=1), 1.0, atol=1e-15), k


# --------------------------------------------------------- 5x5 kernels ------

def _icy5(golden):
    return golden("worlds")["icy_5_0.2"]


def test_backward_5x5(golden):
    g, P = golden("kernels"), _icy5(golden)
    pa = D.local_action_probabilities(P, [24], g["k5_reward"])
    assert np.array_equal(pa, g["k5_lap"])
    pa_x = D.local_action_probabilities(P, [24], g["k5_reward"], rescale=True)
    assert np.array_equal(pa_x, g["k5_lap"]), "pow-2 rescale must be bit-identical where finite"
    # SURVEY 8c KAT1
    assert pa[0, 0] == 0.35662558490748464 and pa[23, 0] == 0.6611946180385522


def test_svf_5x5(golden):
    g, P = golden("kernels"), _icy5(golden)
    d, n = D.expected_svf_from_policy(P, g["k5_p0"], [24], g["k5_lap"])
    assert n == int(g["k5_svf_n"]) == 72
    assert np.array_equal(d, g["k5_svf"])
    assert d.sum() == 17.827970728063327


@pytest.mark.parametrize("gamma,n_lap,n_svf", [(0.7, 521, 497), (0.9, 771, 216)])
def test_causal_5x5(golden, gamma, n_lap, n_svf):
    g, P = golden("kernels"), _icy5(golden)
    pc, n = D.local_causal_action_probabilities(P, [24], g["k5_reward"], gamma)
    assert n == int(g["k5_lcap_%s_n" % gamma]) == n_lap
    assert np.array_equal(pc, g["k5_lcap_%s" % gamma])
    d, n = D.expected_svf_from_policy(P, g["k5_p0"], [24], pc)
    assert n == int(g["k5_csvf_%s_n" % gamma]) == n_svf
    assert np.array_equal(d, g["k5_csvf_%s" % gamma])


def test_causal_terminal_reward_array(golden):
    g, P = golden("kernels"), _icy5(golden)
    pc, n = D.local_causal_action_probabilities(P, g["k5_phi"], g["k5_reward"], 0.8, 1e-6)
    assert n == int(g["k5_lcap_phi_n"])
    assert np.array_equal(pc, g["k5_lcap_phi"])


def test_value_iteration_5x5(golden):
    g, P = golden("kernels"), _icy5(golden)
    v, n = D.value_iteration(P, g["k5_reward"], 0.7)
    assert n == int(g["k5_vi_n"]) == 21
    assert np.array_equal(v, g["k5_vi"])
    v, n = D.value_iteration(P, g["k5_reward"], 0.9, 1e-6)
    assert n == int(g["k5_vi9_n"]) and np.array_equal(v, g["k5_vi9"])


def test_two_terminals(golden):
    g, P = golden("kernels"), _icy5(golden)
    pa = D.local_action_probabilities(P, [24, 4], g["k5b_reward"])
    assert np.array_equal(pa, g["k5b_lap"])
    d, n = D.expected_svf_from_policy(P, g["k5b_p0"], [24, 4], pa, 1e-7)
    assert n == int(g["k5b_svf_n"]) and np.array_equal(d, g["k5b_svf"])


@pytest.mark.parametrize("n", [8, 12])
def test_larger_grids_dense_and_sparse(golden, n):
    g = golden("kernels")
    S = n * n
    P = D.icy_gridworld_table(n, 0.2)
    mdp = SP.icy_gridworld_sparse(n, 0.2)
    pre = "k%d_" % n
    p0 = np.zeros(S); p0[0] = 1.0
    term = [S - 1]

    pa = D.local_action_probabilities(P, term, g[pre + "reward"])
    assert np.array_equal(pa, g[pre + "lap"])
    assert np.array_equal(D.local_action_probabilities(P, term, g[pre + "reward"], rescale=True), pa)
    pa_s = SP.local_action_probabilities(mdp, term, g[pre + "reward"])
    np.testing.assert_allclose(pa_s, pa, rtol=1e-12, atol=0)

    d, k = D.expected_svf_from_policy(P, p0, term, pa)
    assert k == int(g[pre + "svf_n"]) and np.array_equal(d, g[pre + "svf"])
    d_s, k_s = SP.expected_svf_from_policy(mdp, p0, term, pa)
    assert k_s == k
    np.testing.assert_allclose(d_s, d, rtol=1e-12, atol=1e-300)

    pc, k = D.local_causal_action_probabilities(P, term, g[pre + "greward"], 0.9)
    assert k == int(g[pre + "lcap_n"]) and np.array_equal(pc, g[pre + "lcap"])
    pc_s, k_s = SP.local_causal_action_probabilities(mdp, term, g[pre + "greward"], 0.9)
    assert k_s == k
    np.testing.assert_allclose(pc_s, pc, rtol=1e-12, atol=0)

    d, k = D.expected_svf_from_policy(P, p0, term, pc)
    assert k == int(g[pre + "csvf_n"]) and np.array_equal(d, g[pre + "csvf"])

    v, k = D.value_iteration(P, g[pre + "greward"], 0.95, 1e-5)
    assert k == int(g[pre + "vi_n"]) and np.array_equal(v, g[pre + "vi"])
    v_s, k_s = SP.value_iteration(mdp, g[pre + "greward"], 0.95, 1e-5)
    assert k_s == k
    np.testing.assert_allclose(v_s, v, rtol=1e-12, atol=0)


def test_overflow_regime(golden):
    """13x13 with reward == 1: the raw reference overflows to NaN; the
    range-extended oracle stays finite and normalised (SURVEY TL;DR 2)."""
    g = golden("kernels")
    P = D.icy_gridworld_table(13, 0.2)
    with np.errstate(all="ignore"):
        raw = D.local_action_probabilities(P, [168], np.ones(169))
    assert np.isnan(g["k13_lap_overflow"]).all() and np.isnan(raw).all()
    ext = D.local_action_probabilities(P, [168], np.ones(169), rescale=True)
    assert np.isfinite(ext).all()
    np.testing.assert_allclose(ext.sum(axis=1), 1.0, rtol=1e-12)


# ------------------------------------------------------- random MDPs --------

@pytest.mark.parametrize("i", range(6))
def test_random_mdps(golden, i):
    g = golden("random_mdps")
    pre = "r%d_" % i
    P, r, p0 = g[pre + "P"], g[pre + "reward"], g[pre + "p0"]
    term = list(g[pre + "terminal"])
    assert np.array_equal(D.local_action_probabilities(P, term, r), g[pre + "lap"])
    pc, n = D.local_causal_action_probabilities(P, term, r, 0.85)
    assert n == int(g[pre + "lcap_n"]) and np.array_equal(pc, g[pre + "lcap"])
    d, n = D.expected_svf_from_policy(P, p0, term, g[pre + "pol"])
    assert n == int(g[pre + "svf_n"]) and np.array_equal(d, g[pre + "svf"])
    v, n = D.value_iteration(P, r, 0.9, 1e-6)
    assert n == int(g[pre + "vi_n"]) and np.array_equal(v, g[pre + "vi"])
    # sparse restatement on non-grid structure
    mdp = SP.SparseMDP.from_dense(P)
    d_s, n_s = SP.expected_svf_from_policy(mdp, p0, term, g[pre + "pol"])
    np.testing.assert_allclose(d_s, g[pre + "svf"], rtol=1e-11, atol=1e-300)
    assert n_s == int(g[pre + "svf_n"])


# ---------------------------------------------------------- end to end ------

def test_trajectory_statistics(golden):
    g = golden("e2e_5x5")
    tjs = load_trajectories(g)
    assert len(tjs) == 200
    fe = D.feature_expectation_from_trajectories(np.identity(25), tjs)
    assert np.array_equal(fe, g["e_features"])
    assert np.array_equal(D.initial_probabilities_from_trajectories(25, tjs), g["p_initial"])


def test_expert_pipeline(golden):
    """main.py:32-51: value iteration -> weighted stochastic policy."""
    g, k = golden("e2e_5x5"), golden("kernels")
    P = _icy5(golden)
    v, _ = D.value_iteration(P, k["k5_reward"], 0.7)
    assert np.array_equal(v, g["expert_value"])
    pol = D.stochastic_policy_from_value(5, 4, v, w=lambda x: x ** 5)
    assert np.array_equal(pol, g["expert_policy"])


def test_irl_5x5_end_to_end(golden):
    g = golden("e2e_5x5")
    P = _icy5(golden)
    F = np.identity(25)
    opt = D.ExpSgaPort(lr=D.linear_decay(lr0=0.2))
    r, n, _ = D.irl(P, F, [24], g["e_features"], g["p_initial"], opt, np.ones(25))
    assert n == int(g["irl_steps"]) == 375
    assert np.array_equal(r, g["irl_reward"])


@pytest.mark.parametrize("gamma,steps", [(0.9, 382)])
def test_irl_causal_5x5_end_to_end(golden, gamma, steps):
    g = golden("e2e_5x5")
    P = _icy5(golden)
    F = np.identity(25)
    opt = D.ExpSgaPort(lr=D.linear_decay(lr0=0.2))
    r, n, _ = D.irl_causal(P, F, [24], g["e_features"], g["p_initial"], opt, np.ones(25), gamma)
    assert n == int(g["irl_causal_%s_steps" % gamma]) == steps
    assert np.array_equal(r, g["irl_causal_%s_reward" % gamma])


def test_irl_nan_exit(golden):
    """NaN ends every loop: Sga + coordinate features overflows the raw pass."""
    g = golden("e2e_5x5")
    P = _icy5(golden)
    F = g["coord_features"]
    tjs = load_trajectories(g)
    ef = D.feature_expectation_from_trajectories(F, tjs)
    with np.errstate(all="ignore"):
        r, n, _ = D.irl(P, F, [24], ef, g["p_initial"], D.SgaPort(0.02), np.full(5, -0.3), eps=1e-3)
    assert n == int(g["irl_sga_coord_nan_steps"])
    assert np.isnan(r).all() and np.isnan(g["irl_sga_coord_nan_reward"]).all()
