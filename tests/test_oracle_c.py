"""The plain-C restatement (oracle/c/irl_oracle.c) against the numpy restatement, which is pinned
bit-for-bit to the unmodified reference.  CPU only."""
import numpy as np
import pytest

from oracle import c_port as C
from oracle import dense_port as D
from oracle import sparse_port as SP


def test_c_oracle_5x5(golden):
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    sidx, sp = C.ell_from_dense(P)
    pa = C.backward(sidx, sp, [24], g["k5_reward"])
    np.testing.assert_allclose(pa, g["k5_lap"], rtol=1e-13)
    np.testing.assert_allclose(C.backward(sidx, sp, [24], g["k5_reward"], rescale=True), g["k5_lap"], rtol=1e-13)
    d, n = C.svf(sidx, sp, g["k5_p0"], [24], g["k5_lap"])
    assert n == int(g["k5_svf_n"]) == 72
    np.testing.assert_allclose(d, g["k5_svf"], rtol=1e-13)
    for gamma in (0.7, 0.9):
        phi = D.terminal_reward([24], 25)
        pc, _, n = C.soft_vi(sidx, sp, phi, g["k5_reward"], gamma)
        assert n == int(g["k5_lcap_%s_n" % gamma])
        np.testing.assert_allclose(pc, g["k5_lcap_%s" % gamma], rtol=1e-12)
    pc, _, n = C.soft_vi(sidx, sp, g["k5_phi"], g["k5_reward"], 0.8, 1e-6)
    assert n == int(g["k5_lcap_phi_n"])
    np.testing.assert_allclose(pc, g["k5_lcap_phi"], rtol=1e-12)
    v, n = C.value_iteration(sidx, sp, g["k5_reward"], 0.7)
    assert n == 21
    np.testing.assert_allclose(v, g["k5_vi"], rtol=1e-14)


@pytest.mark.parametrize("i", range(6))
def test_c_oracle_random_mdps(golden, i):
    g, pre = golden("random_mdps"), "r%d_" % i
    P, r, p0 = g[pre + "P"], g[pre + "reward"], g[pre + "p0"]
    term = list(g[pre + "terminal"])
    sidx, sp = C.ell_from_dense(P)
    with np.errstate(invalid="ignore"):
        np.testing.assert_allclose(C.backward(sidx, sp, term, r), g[pre + "lap"], rtol=1e-12)
    pc, _, n = C.soft_vi(sidx, sp, D.terminal_reward(term, P.shape[0]), r, 0.85)
    assert n == int(g[pre + "lcap_n"])
    np.testing.assert_allclose(pc, g[pre + "lcap"], rtol=1e-12)
    d, n = C.svf(sidx, sp, p0, term, g[pre + "pol"])
    assert n == int(g[pre + "svf_n"])
    np.testing.assert_allclose(d, g[pre + "svf"], rtol=1e-12, atol=1e-300)
    v, n = C.value_iteration(sidx, sp, r, 0.9, 1e-6)
    assert n == int(g[pre + "vi_n"])
    np.testing.assert_allclose(v, g[pre + "vi"], rtol=1e-13)


def test_c_oracle_batch_and_sparse_tables(golden):
    g = golden("kernels")
    n = 12
    S = n * n
    mdps = [SP.icy_gridworld_sparse(n, p) for p in (0.2, 0.3)]
    tabs = [C.ell_from_sparse(m) for m in mdps]
    sidx = np.stack([t[0] for t in tabs]); sp = np.stack([t[1] for t in tabs])
    # same table as from the dense route
    dsidx, dsp = C.ell_from_dense(D.icy_gridworld_table(n, 0.2))
    assert np.array_equal(dsidx, tabs[0][0]) and np.array_equal(dsp, tabs[0][1])
    p0 = np.zeros(S); p0[0] = 1.0
    rewards = np.stack([g["k12_reward"], g["k12_reward"][::-1].copy()])
    out, nn = C.batch_maxent_step(sidx, sp, [S - 1], p0, rewards)
    assert nn[0] == int(g["k12_svf_n"])
    np.testing.assert_allclose(out[0], g["k12_svf"], rtol=1e-12)
    ref, n_ref = D.compute_expected_svf(D.icy_gridworld_table(n, 0.3), p0, [S - 1], rewards[1])
    assert nn[1] == n_ref
    np.testing.assert_allclose(out[1], ref, rtol=1e-12)


@pytest.mark.parametrize("n,p", [(5, 0.2), (24, 0.3)])
def test_threaded_forward_oracle_is_bitwise_the_serial_loop(n, p):
    """oracle_svf_mt (OpenMP gather over targets, used for the full-size C3 checks) adds every target's
    contributions in the order of the serial scatter: same bits, same sweep count."""
    S = n * n
    sidx, sp = C.ell_from_sparse(SP.icy_gridworld_sparse(n, p))
    rng = np.random.default_rng(n)
    r = -0.1 + 0.05 * rng.standard_normal(S); r[S - 1] = 1.0
    pol, _, _ = C.soft_vi(sidx, sp, D.terminal_reward([S - 1], S), r, 0.9)
    p0 = np.zeros(S); p0[0] = 0.5; p0[S // 2] = 0.5
    for ms in (0, 37):
        d0, n0 = C.svf(sidx, sp, p0, [S - 1], pol, 1e-5, max_sweeps=ms, threads=False)
        d1, n1 = C.svf(sidx, sp, p0, [S - 1], pol, 1e-5, max_sweeps=ms, threads=True)
        assert n0 == n1 and np.array_equal(d0, d1)
    bad = pol.copy(); bad[3, 1] = np.nan                      # NaN ends both loops on the same sweep
    d0, n0 = C.svf(sidx, sp, p0, [S - 1], bad, 1e-5, threads=False)
    d1, n1 = C.svf(sidx, sp, p0, [S - 1], bad, 1e-5, threads=True)
    assert n0 == n1 and np.array_equal(np.isnan(d0), np.isnan(d1))
