"""Parity of the CUDA path (through the C ABI) against the oracle and the
reference-generated golden fixtures.  Needs a B200: `pytest -m gpu`.

Tolerance (BASELINE.json north_star): float64 mode within 1e-10 relative on SVF,
policy and learned reward, identical iteration counts.  Indices and table values
are compared bit-exactly.
"""
import os

import numpy as np
import pytest

import _irlb200 as E
import gridworld as W
import maxent as M
import optimizer as O
import solver as S
import trajectory as T

from oracle import dense_port as D
from oracle import sparse_port as SP
from test_oracle_golden import load_trajectories

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def close(a, b, rtol=RTOL, atol=1e-300):
    a = a.cpu().numpy() if E.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def counts():
    return E.last_info.counts()


@pytest.fixture(params=["auto", "streamed", "grid", "grid_streamed"])
def variant(request, monkeypatch):
    """Run the same parity checks through every execution variant of the kernels."""
    v = request.param
    if "streamed" in v:
        monkeypatch.setenv("IRLB200_FORCE_STREAMED", "1")
    return E.MODE_GRID if v.startswith("grid") else E.MODE_AUTO


# ------------------------------------------------------------------ tables ---

def expected_tables(P):
    """Reference structure straight from the dense table with numpy."""
    S, _, A = P.shape
    nz = (P != 0).any(axis=2)
    succ = [np.nonzero(nz[s])[0] for s in range(S)]
    pred = [np.nonzero(nz[:, s])[0] for s in range(S)]
    return succ, pred


def check_tables(t, P):
    S, _, A = P.shape
    succ, pred = expected_tables(P)
    si, sp = t.succ_idx[0].cpu().numpy(), t.succ_p[0].cpu().numpy()
    pi, pp = t.pred_idx[0].cpu().numpy(), t.pred_p[0].cpu().numpy()
    assert si.shape == (t.Ks, S) and sp.shape == (A, t.Ks, S)
    for s in range(S):
        n = len(succ[s])
        assert np.array_equal(si[:n, s], succ[s]), "successor indices must be bit-exact"
        assert (si[n:, s] == s).all() and (sp[:, n:, s] == 0).all()
        assert np.array_equal(sp[:, :n, s], P[s, succ[s], :].T)
        m = len(pred[s])
        assert np.array_equal(pi[:m, s], pred[s]), "predecessor indices must be bit-exact"
        assert (pi[m:, s] == s).all() and (pp[:, m:, s] == 0).all()
        assert np.array_equal(pp[:, :m, s], P[pred[s], s, :].T)


@pytest.mark.parametrize("key", ["icy_1_0.2", "icy_2_0.2", "icy_3_0.35", "icy_5_0.2", "icy_8_0.2", "grid_5", "grid_2"])
def test_compress_dense_worlds(golden, key):
    P = golden("worlds")[key]
    t = E.compress_dense(P)
    check_tables(t, P)
    succ, pred = expected_tables(P)
    assert t.k_discovered == (max(map(len, succ)), max(map(len, pred)))


@pytest.mark.parametrize("i", range(6))
def test_compress_dense_random(golden, i):
    P = golden("random_mdps")["r%d_P" % i]
    check_tables(E.compress_dense(P), P)


@pytest.mark.parametrize("n,p,icy", [(1, 0.2, True), (2, 0.2, True), (5, 0.2, True), (8, 0.35, True),
                                     (5, 0.0, False), (17, 0.1, True)])
def test_direct_world_tables_equal_compressed_dense(golden, n, p, icy):
    """gridworld_tables (no dense detour) == compress(dense), bit for bit; and the
    device-built dense table == the reference's table."""
    Pd = E.gridworld_dense(n, p, icy)
    key = ("icy_%d_%s" % (n, p)) if icy else "grid_%d" % n
    g = golden("worlds")
    if key in g.files:
        assert np.array_equal(Pd.cpu().numpy(), g[key])
    else:
        ref = D.icy_gridworld_table(n, p) if icy else D.gridworld_table(n)
        assert np.array_equal(Pd.cpu().numpy(), ref)
    a = E.compress_dense(Pd)
    b = E.gridworld_tables(n, p, icy=icy)
    if a.Ks == 5 and a.Kp == 5:
        for x, y in ((a.succ_idx, b.succ_idx), (a.succ_p, b.succ_p), (a.pred_idx, b.pred_idx), (a.pred_p, b.pred_p)):
            assert (x == y).all()
    check_tables(b, Pd.cpu().numpy())


def test_batched_world_tables():
    ps = [0.1, 0.2, 0.3]
    t = E.gridworld_tables(6, ps)
    assert t.n_tables == 3
    for b, p in enumerate(ps):
        check_tables(t.select(b), D.icy_gridworld_table(6, p))


# ------------------------------------------------------- per-kernel parity ---

def test_backward_5x5(golden, variant):
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    t = E.compress_dense(P)
    pol = E.backward(t, E.terminal_mask([24], 25), g["k5_reward"], mode=variant)
    close(pol[0], g["k5_lap"])
    pol = E.backward(t, E.terminal_mask([24, 4], 25), g["k5b_reward"], mode=variant)
    close(pol[0], g["k5b_lap"])


def test_svf_5x5(golden, variant):
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    t = E.compress_dense(P)
    d = E.svf(t, g["k5_p0"], E.terminal_mask([24], 25), g["k5_lap"], 1e-5, mode=variant)
    assert counts()[0] == int(g["k5_svf_n"]) == 72
    close(d[0], g["k5_svf"])
    d = E.svf(t, g["k5b_p0"], E.terminal_mask([24, 4], 25), g["k5b_lap"], 1e-7, mode=variant)
    assert counts()[0] == int(g["k5b_svf_n"])
    close(d[0], g["k5b_svf"])


@pytest.mark.parametrize("gamma", [0.7, 0.9])
def test_soft_vi_5x5(golden, variant, gamma):
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    t = E.compress_dense(P)
    pol = E.soft_vi(t, E.terminal_phi([24], 25), g["k5_reward"], gamma, 1e-5, mode=variant)
    assert counts()[0] == int(g["k5_lcap_%s_n" % gamma])
    close(pol[0], g["k5_lcap_%s" % gamma])
    d = E.svf(t, g["k5_p0"], E.terminal_mask([24], 25), pol[0], 1e-5, mode=variant)
    assert counts()[0] == int(g["k5_csvf_%s_n" % gamma])
    close(d[0], g["k5_csvf_%s" % gamma])


def test_soft_vi_terminal_reward_array(golden, variant):
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    t = E.compress_dense(P)
    pol = E.soft_vi(t, E.terminal_phi(g["k5_phi"], 25), g["k5_reward"], 0.8, 1e-6, mode=variant)
    assert counts()[0] == int(g["k5_lcap_phi_n"])
    close(pol[0], g["k5_lcap_phi"])


def test_value_iteration_5x5(golden, variant):
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    t = E.compress_dense(P)
    v = E.value_iteration(t, g["k5_reward"], 0.7, 1e-3, mode=variant)
    assert counts()[0] == 21
    close(v[0], g["k5_vi"])
    v = E.value_iteration(t, g["k5_reward"], 0.9, 1e-6, mode=variant)
    assert counts()[0] == int(g["k5_vi9_n"])
    close(v[0], g["k5_vi9"])


@pytest.mark.parametrize("n", [8, 12])
def test_larger_grids(golden, variant, n):
    g = golden("kernels")
    S, pre = n * n, "k%d_" % n
    t = E.gridworld_tables(n, 0.2)
    p0 = np.zeros(S); p0[0] = 1.0
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    pol = E.backward(t, mask, g[pre + "reward"], mode=variant)
    close(pol[0], g[pre + "lap"])
    d = E.svf(t, p0, mask, g[pre + "lap"], mode=variant)
    assert counts()[0] == int(g[pre + "svf_n"])
    close(d[0], g[pre + "svf"])
    pol = E.soft_vi(t, phi, g[pre + "greward"], 0.9, mode=variant)
    assert counts()[0] == int(g[pre + "lcap_n"])
    close(pol[0], g[pre + "lcap"])
    d = E.svf(t, p0, mask, g[pre + "lcap"], mode=variant)
    assert counts()[0] == int(g[pre + "csvf_n"])
    close(d[0], g[pre + "csvf"])
    v = E.value_iteration(t, g[pre + "greward"], 0.95, 1e-5, mode=variant)
    assert counts()[0] == int(g[pre + "vi_n"])
    close(v[0], g[pre + "vi"])


@pytest.mark.parametrize("i", range(6))
def test_random_mdps(golden, variant, i):
    """Non-grid MDPs: run-time A and K, ragged rows."""
    g, pre = golden("random_mdps"), "r%d_" % i
    P = g[pre + "P"]
    S = P.shape[0]
    term = list(g[pre + "terminal"])
    t = E.compress_dense(P)
    mask, phi = E.terminal_mask(term, S), E.terminal_phi(term, S)
    close(E.backward(t, mask, g[pre + "reward"], mode=variant)[0], g[pre + "lap"])
    pol = E.soft_vi(t, phi, g[pre + "reward"], 0.85, mode=variant)
    assert counts()[0] == int(g[pre + "lcap_n"])
    close(pol[0], g[pre + "lcap"])
    d = E.svf(t, g[pre + "p0"], mask, g[pre + "pol"], mode=variant)
    assert counts()[0] == int(g[pre + "svf_n"])
    close(d[0], g[pre + "svf"])
    v = E.value_iteration(t, g[pre + "reward"], 0.9, 1e-6, mode=variant)
    assert counts()[0] == int(g[pre + "vi_n"])
    close(v[0], g[pre + "vi"])


def test_overflow_regime_is_range_extended(golden):
    """13x13, reward == 1: the raw reference returns NaN; the product equals the
    range-extended oracle (SURVEY TL;DR 2)."""
    P = D.icy_gridworld_table(13, 0.2)
    ext = D.local_action_probabilities(P, [168], np.ones(169), rescale=True)
    pol = M.local_action_probabilities(P, [168], np.ones(169))
    assert np.isfinite(pol).all()
    close(pol, ext)
    # underflow side
    ext = D.local_action_probabilities(P, [168], np.full(169, -6.0), rescale=True)
    close(M.local_action_probabilities(P, [168], np.full(169, -6.0)), ext)


def test_guards_and_status(golden):
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    t = E.compress_dense(P)
    E.svf(t, g["k5_p0"], E.terminal_mask([24], 25), g["k5_lap"], 1e-5, max_sweeps=10)
    assert counts()[0] == 10 and E.last_info.stati()[0] == E.ST_MAXSWEEPS
    # NaN ends the loop in the sweep it appears (reference: `while delta > eps`)
    pol = np.array(g["k5_lap"]); pol[3, 1] = np.nan
    E.svf(t, g["k5_p0"], E.terminal_mask([24], 25), pol, 1e-5)
    assert E.last_info.stati()[0] == E.ST_NONFINITE
    d_ref, n_ref = D.expected_svf_from_policy(P, g["k5_p0"], [24], pol)
    # the reference's dense 0*NaN poisons every state in the first sweep; the sparse gather needs a
    # few hops, and the register-resident kernel looks for a non-finite iterate every 16 sweeps
    assert counts()[0] <= n_ref + 16


@pytest.mark.parametrize("n", [8, 32])
def test_nan_exit_sweep_against_the_oracle(n, monkeypatch):
    """`while delta > eps` ends on NaN (maxent.py:108).  Against the sparse restatement's NaN-exit sweep: the
    generic kernels (template, streamed, cooperative grid) and the slab kernels stop on exactly that sweep; the
    hand-tuned tiled / fast forward kernels look for a non-finite iterate every 16 sweeps and therefore stop on
    the next multiple of 16 -- never earlier, at most 15 sweeps later (documented in DESIGN.md section 2)."""
    from oracle import c_port as C
    S = n * n
    sidx, sp = C.ell_from_sparse(SP.icy_gridworld_sparse(n, 0.2))
    r = np.full(S, -0.1); r[S - 1] = 1.0
    phi = np.full(S, -np.inf); phi[S - 1] = 0.0
    pol, _, _ = C.soft_vi(sidx, sp, phi, r, 0.9)
    p0 = np.zeros(S); p0[0] = 1.0
    far = S - 2                                  # poison a state the start mass needs several sweeps to reach
    pol[far, 2] = np.nan
    _, n_ref = C.svf(sidx, sp, p0, [S - 1], pol, 1e-5)
    assert n_ref >= 1
    t = E.gridworld_tables(n, 0.2)
    mask = E.terminal_mask([S - 1], S)
    E.svf(t, p0, mask, pol, 1e-5)                                        # tiled kernel
    n_tiled, st = int(counts()[0]), int(E.last_info.stati()[0])
    assert st == E.ST_NONFINITE and n_ref <= n_tiled <= n_ref + 15 and n_tiled % 16 == 0
    monkeypatch.setenv("IRLB200_SVF_TILE", "0")
    monkeypatch.setenv("IRLB200_FORCE_STREAMED", "1")                    # generic template kernel
    E.svf(t, p0, mask, pol, 1e-5)
    assert int(counts()[0]) == n_ref and int(E.last_info.stati()[0]) == E.ST_NONFINITE
    E.svf(t, p0, mask, pol, 1e-5, mode=E.MODE_GRID)                      # cooperative grid
    assert int(counts()[0]) == n_ref and int(E.last_info.stati()[0]) == E.ST_NONFINITE
    import slab
    g = slab.PeerSlabGrid(n, 0.2, flow=True, chunk_sweeps=5)             # dataflow slab kernel: exact, via replay
    try:
        g.svf(p0, [S - 1], pol, 1e-5)
        assert (g.last_n_iter, g.last_status) == (n_ref, E.ST_NONFINITE)
    finally:
        g.close()


@pytest.mark.timeout(120)
def test_slab_flow_kernel_aborts_instead_of_hanging_when_a_peer_is_missing():
    """Failure detection of the multi-GPU dataflow kernel, on ONE GPU: rank 0 of a world of two is
    launched while "rank 1" (a second local block standing in for the peer's memory) never runs.  The CTAs of the
    boundary row wait for mailbox values that never arrive; after `timeout_s` they raise the abort word, every
    spin loop sees it, all CTAs meet at the chunk barrier and the launch returns IRLB200_ST_ABORTED -- no hang."""
    import ctypes
    import time
    import torch
    n = 64
    S, half = n * n, n * n // 2
    lib = E.load_library()
    E.require_cuda()
    dev = E._dev()
    blocks = []
    try:
        for _ in range(2):
            p = ctypes.c_void_p()
            E._check(lib.irlb200_peer_alloc(lib.irlb200_slab_flow_block_bytes(S, n), ctypes.byref(p)))
            E._check(lib.irlb200_slab_flow_reset(p, S, n, E._stream()))
            blocks.append(p)
        arr = (ctypes.c_void_p * 2)(blocks[0].value, blocks[1].value)
        K, A = 4, 4
        t = dict(idx=torch.empty((K, half), dtype=torch.int32, device=dev),
                 p=torch.empty((A, K, half), dtype=torch.float64, device=dev),
                 pidx=torch.empty((K, half), dtype=torch.int32, device=dev),
                 pp=torch.empty((A, K, half), dtype=torch.float64, device=dev))
        E._check(lib.irlb200_gridworld_tables_range_k(n, 1, 0.2, 0, half, K, E._ptr(t["idx"]), E._ptr(t["p"]),
                                                      E._ptr(t["pidx"]), E._ptr(t["pp"]), E._stream()))
        r = torch.full((half,), -0.1, dtype=torch.float64, device=dev)
        out = torch.empty(half, dtype=torch.float64, device=dev)
        n_iter = torch.zeros(1, dtype=torch.int32, device=dev)
        status = torch.full((1,), -7, dtype=torch.int32, device=dev)
        work = torch.empty(int(lib.irlb200_slab_flow_work_bytes(half)), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        E._check(lib.irlb200_slab_flow(2, 0, 2, arr, S, 0, half, n, A, K, E._ptr(t["idx"]), E._ptr(t["p"]), E._ptr(r),
                                       None, None, None, None, 0.9, 1e-4, 100000, 0, E._ptr(out), None, E._ptr(n_iter),
                                       E._ptr(status), 0.5, 0, E._ptr(work), work.numel(), E._stream()))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert int(status.item()) == E.ST_ABORTED, (int(status.item()), int(n_iter.item()))
        assert 0.4 < dt < 10.0, dt
    finally:
        torch.cuda.synchronize()
        for p in blocks:
            lib.irlb200_peer_free(p)


# ------------------------------------------------------------- public API ---

def test_module_api_numpy_roundtrip(golden):
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    pa = M.local_action_probabilities(P, [24], g["k5_reward"])
    assert isinstance(pa, np.ndarray) and pa.shape == (25, 4)
    close(pa, g["k5_lap"])
    close(M.expected_svf_from_policy(P, g["k5_p0"], [24], pa), g["k5_svf"])
    close(M.compute_expected_svf(P, g["k5_p0"], [24], g["k5_reward"]), g["k5_svf"])
    close(M.local_causal_action_probabilities(P, [24], g["k5_reward"], 0.9), g["k5_lcap_0.9"])
    close(M.compute_expected_causal_svf(P, g["k5_p0"], [24], g["k5_reward"], 0.7), g["k5_csvf_0.7"])
    v = S.value_iteration(P, g["k5_reward"], 0.7)
    assert isinstance(v, np.ndarray) and v.shape == (25,)
    close(v, g["k5_vi"])
    close(M.softmax(np.array([1.0, -np.inf]), np.array([2.0, 3.0])), D.softmax(np.array([1.0, -np.inf]), np.array([2.0, 3.0])))


def test_module_api_tensor_in_tensor_out(golden):
    import torch
    g, P = golden("kernels"), golden("worlds")["icy_5_0.2"]
    Pt = torch.as_tensor(P).cuda()
    r = torch.as_tensor(g["k5_reward"]).cuda()
    pa = M.local_action_probabilities(Pt, [24], r)
    assert isinstance(pa, torch.Tensor) and pa.is_cuda
    close(pa, g["k5_lap"])
    d = M.compute_expected_svf(W.IcyGridWorld(5).tables(), torch.as_tensor(g["k5_p0"]).cuda(), [24], r)
    assert d.is_cuda
    close(d, g["k5_svf"])
    v = S.value_iteration(Pt, r, 0.7)
    assert v.is_cuda
    close(v, g["k5_vi"])


def test_in_place_edit_of_a_dense_table_is_not_served_from_the_cache():
    """The reference re-reads p_transition on every call; the table cache must notice an in-place edit of a
    table larger than 1 MiB (16x16: 2 MB) and of a CUDA tensor."""
    n = 16
    S = n * n
    P = D.icy_gridworld_table(n, 0.2)
    assert P.nbytes > (1 << 20)
    r = -0.1 * np.ones(S); r[S - 1] = 1.0
    import solver as SV
    v0 = SV.value_iteration(P, r, 0.9, 1e-6)
    P[:] = D.icy_gridworld_table(n, 0.35)                # same object, new contents
    v1 = SV.value_iteration(P, r, 0.9, 1e-6)
    ref, _ = D.value_iteration(P, r, 0.9, 1e-6)
    close(v1, ref)
    assert np.max(np.abs(v1 - v0)) > 1e-6
    Pt = E.to_device(D.icy_gridworld_table(n, 0.2))
    w0 = SV.value_iteration(Pt, E.to_device(r), 0.9, 1e-6)
    close(w0, v0)
    Pt.copy_(E.to_device(P))                             # in place on the device
    w1 = SV.value_iteration(Pt, E.to_device(r), 0.9, 1e-6)
    close(w1, ref)


def test_fused_step_matches_split(golden):
    g = golden("kernels")
    t = E.gridworld_tables(8, 0.2)
    p0 = np.zeros(64); p0[0] = 1.0
    mask, phi = E.terminal_mask([63], 64), E.terminal_phi([63], 64)
    for fused in (True, False):
        d, grad, pol = E.expected_svf(t, p0, mask, g["k8_reward"], causal=False, e_features=p0,
                                      want_policy=True, fused=fused)
        close(d[0], g["k8_svf"])
        close(pol[0], g["k8_lap"])
        close(grad[0], p0 - g["k8_svf"], atol=1e-12)
        assert E.last_info.counts()[0, 1] == int(g["k8_svf_n"])
        d, _, pol = E.expected_svf(t, p0, mask, g["k8_greward"], causal=True, phi=phi, discount=0.9,
                                   want_policy=True, fused=fused)
        close(d[0], g["k8_csvf"])
        close(pol[0], g["k8_lcap"])
        assert tuple(E.last_info.counts()[0]) == (int(g["k8_lcap_n"]), int(g["k8_csvf_n"]))


def test_expert_pipeline_and_solver_helpers(golden):
    g, k = golden("e2e_5x5"), golden("kernels")
    world = W.IcyGridWorld(5, 0.2)
    value = S.value_iteration(world.p_transition, k["k5_reward"], 0.7)
    close(value, g["expert_value"])
    pol = S.stochastic_policy_from_value(world, g["expert_value"], w=lambda x: x ** 5)
    assert np.array_equal(pol, g["expert_policy"])
    greedy = S.optimal_policy(world, k["k5_reward"], 0.7)
    ref = np.array([np.argmax([g["expert_value"][world.state_index_transition(s, a)] for a in range(4)])
                    for s in range(25)])
    assert np.array_equal(greedy, ref)
    sv = S.stochastic_value_iteration(world.p_transition, k["k5_reward"], 0.7)
    assert sv.shape == (25,) and np.isfinite(sv).all()


# ------------------------------------------------------------- end to end ---

class _Count:
    def __init__(self, inner):
        self.inner, self.n = inner, 0

    def reset(self, p):
        self.inner.reset(p)

    def step(self, g, *a, **k):
        self.n += 1
        return self.inner.step(g, *a, **k)


def test_irl_5x5_like_main(golden):
    """The call sequence of the reference's main.py (main.py:54-72) on the fixture
    trajectories: 375 outer steps, learned reward to 1e-10."""
    g = golden("e2e_5x5")
    world = W.IcyGridWorld(size=5, p_slip=0.2)
    tjs = load_trajectories(g)
    opt = _Count(O.ExpSga(lr=O.linear_decay(lr0=0.2)))
    r = M.irl(world.p_transition, W.state_features(world), [24], tjs, opt, O.Constant(1.0))
    assert isinstance(r, np.ndarray)
    assert opt.n == int(g["irl_steps"]) == 375
    close(r, g["irl_reward"])


@pytest.mark.parametrize("gamma", [0.7, 0.9])
def test_irl_causal_5x5_like_main(golden, gamma):
    g = golden("e2e_5x5")
    world = W.IcyGridWorld(size=5, p_slip=0.2)
    tjs = load_trajectories(g)
    opt = _Count(O.ExpSga(lr=O.linear_decay(lr0=0.2)))
    r = M.irl_causal(world.p_transition, W.state_features(world), [24], tjs, opt, O.Constant(1.0), gamma)
    assert opt.n == int(g["irl_causal_%s_steps" % gamma])
    close(r, g["irl_causal_%s_reward" % gamma])


def test_the_reference_main_py_runs_unchanged_on_the_gpu(golden):
    """BASELINE north_star: `src/main.py` runs unchanged against the engine.  tests/golden/reference_main_py.fixture
    is a byte-identical copy of the reference's driver (the one verbatim file in this repository: the CALLER that
    must run unmodified; sha256 pinned here and checked against /root/reference/src/main.py where that exists,
    MIT licence beside it).  It is launched through scripts/run_reference_main.py in a fresh interpreter with
    np.random.seed(0); the engine's optimizer classes count the outer steps: 375 (irl) and 419 (irl_causal,
    gamma = 0.7 as in main.py:75) -- the reference's own counts (SURVEY 8c, KAT5 / tests/golden/e2e_5x5.npz)."""
    import hashlib
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fixture = os.path.join(root, "tests", "golden", "reference_main_py.fixture")
    digest = hashlib.sha256(open(fixture, "rb").read()).hexdigest()
    assert digest == "e0f6c933ba28e01809804ddef95d3bdab221edf6da25bc4e1b890b25dbca2cb1"
    ref = "/root/reference/src/main.py"
    if os.path.exists(ref):
        assert hashlib.sha256(open(ref, "rb").read()).hexdigest() == digest
    env = dict(os.environ, IRLB200_MAIN_SEED="0", IRLB200_MAIN_REPORT="1")
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "run_reference_main.py"), fixture],
                       capture_output=True, text=True, timeout=280, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("IRLB200_MAIN_OUTER_STEPS")]
    assert line, p.stdout[-2000:]
    steps = [int(v) for v in line[-1].split()[1:]]
    g = golden("e2e_5x5")
    assert steps == [int(g["irl_steps"]), int(g["irl_causal_0.7_steps"])] == [375, 419]


@pytest.mark.parametrize("causal,gamma", [(False, None), (True, 0.7), (True, 0.9)])
@pytest.mark.parametrize("opt", ["expsga_linear", "sga_const", "expsga_power"])
def test_device_side_outer_loop_equals_the_host_loop(golden, causal, gamma, opt, monkeypatch):
    """irlb200_irl_small (the whole `while delta > eps` loop in one launch, host-evaluated learning rates)
    against the generic host loop: bitwise the same reward, the same number of outer steps, the optimizer
    object left in the same state -- and the reference's own step counts for main.py's settings."""
    g, P = golden("e2e_5x5"), golden("worlds")["icy_5_0.2"]
    tjs = load_trajectories(g)
    F = W.state_features(W.IcyGridWorld(5, 0.2))

    def make():
        if opt == "expsga_linear":
            return O.ExpSga(lr=O.linear_decay(lr0=0.2))
        if opt == "sga_const":
            return O.Sga(lr=0.05)
        return O.ExpSga(lr=O.power_decay(lr0=0.3))

    def run(o):
        if causal:
            return M.irl_causal(P, F, [24], tjs, o, O.Constant(1.0), gamma)
        return M.irl(P, F, [24], tjs, o, O.Constant(1.0))

    monkeypatch.setattr(M, "DEVICE_LOOP_CHUNK", 100)             # several launches: the chunk hand-over is exercised
    dev_opt = make()
    E.launch_log = []
    r_dev = run(dev_opt)
    names = [nm for nm, _, _ in E.launch_log]
    E.launch_log = None
    assert names and set(names) == {"irl_small"}, names
    host_opt = _Count(make())                                    # a wrapper: keeps the generic host loop
    r_host = run(host_opt)
    assert np.array_equal(r_dev, r_host, equal_nan=True)
    assert dev_opt.k == host_opt.n == host_opt.inner.k
    assert len(names) == (dev_opt.k + 99) // 100
    if opt == "expsga_linear":
        key = "irl_steps" if not causal else "irl_causal_%s_steps" % gamma
        assert dev_opt.k == int(g[key])
    monkeypatch.setenv("IRLB200_DEVICE_LOOP", "0")               # the switch
    E.launch_log = []
    r_off = run(make())
    assert "irl_small" not in [nm for nm, _, _ in E.launch_log] and np.array_equal(r_off, r_dev, equal_nan=True)
    E.launch_log = None


def test_irl_other_optimizers_and_dense_features(golden):
    g = golden("e2e_5x5")
    world = W.IcyGridWorld(size=5, p_slip=0.2)
    tjs = load_trajectories(g)
    feats = W.coordinate_features(world)
    opt = _Count(O.ExpSga(lr=O.linear_decay(lr0=0.1)))
    r = M.irl(world.p_transition, feats, [24], tjs, opt, O.Constant(0.2), eps=1e-3)
    assert opt.n == int(g["irl_expsga_coord_steps"])
    close(r, g["irl_expsga_coord_reward"])
    features = W.state_features(world)
    opt = _Count(O.ExpSga(lr=O.power_decay(lr0=0.3)).normalize_grad())
    r = M.irl_causal(world.p_transition, features, [24], tjs, opt, O.Constant(1.0), 0.8, eps=1e-3)
    assert opt.n == int(g["irl_causal_ng_steps"])
    close(r, g["irl_causal_ng_reward"])
    opt = _Count(O.Sga(lr=O.exponential_decay(lr0=0.1, decay_rate=0.05)).normalize_grad())
    r = M.irl_causal(world.p_transition, features, [24], tjs, opt, O.Constant(0.2), 0.8, eps=1e-3)
    assert opt.n == int(g["irl_causal_sga_steps"])
    close(r, g["irl_causal_sga_reward"])


# -------------------------------------------------- batched + larger sizes ---

def test_batch_equals_loop():
    """B independent worlds in one launch == B single launches of the same kernels (bitwise);
    fused and split launches agree to rounding (the split path uses the merged-weight backward
    sweeps, the fused one the per-action form)."""
    n, B = 12, 7
    S = n * n
    ps = 0.1 + 0.2 * np.arange(B) / B
    tabs = E.gridworld_tables(n, ps)
    rng = np.random.default_rng(5)
    rewards = -np.log(4.0) + 0.1 * rng.standard_normal((B, S))
    p0 = np.zeros(S); p0[0] = 1.0
    for causal in (False, True):
        for fused in (True, False):
            d, _ = M.compute_expected_svf_batch(tabs, p0, [S - 1], rewards, causal=causal, discount=0.9, fused=fused)
            nb = E.last_info.counts()
            for b in range(B):
                d1, _ = M.compute_expected_svf_batch(tabs.select(b), p0, [S - 1], rewards[b], causal=causal,
                                                     discount=0.9, fused=fused)
                assert (d[b] == d1[0]).all()
                assert (nb[b] == E.last_info.counts()[0]).all()
                d2, _ = M.compute_expected_svf_batch(tabs.select(b), p0, [S - 1], rewards[b], causal=causal,
                                                     discount=0.9, fused=not fused)
                close(d[b], d2[0].cpu().numpy(), rtol=1e-11)
                assert (nb[b] == E.last_info.counts()[0]).all()
    # one of them against the oracle
    mdp = SP.icy_gridworld_sparse(n, ps[3])
    pa, n_lap = SP.local_causal_action_probabilities(mdp, [S - 1], rewards[3], 0.9)
    dref, n_svf = SP.expected_svf_from_policy(mdp, p0, [S - 1], pa)
    close(d[3], dref)
    assert tuple(nb[3]) == (n_lap, n_svf)


def test_shared_table_batch():
    n, B = 8, 5
    S = n * n
    tabs = E.gridworld_tables(n, 0.2)
    rng = np.random.default_rng(11)
    rewards = -np.log(4.0) + 0.1 * rng.standard_normal((B, S))
    p0 = np.zeros(S); p0[0] = 1.0
    d, _ = M.compute_expected_svf_batch(tabs, p0, [S - 1], rewards)
    P = D.icy_gridworld_table(n, 0.2)
    for b in (0, 4):
        ref, _ = D.compute_expected_svf(P, p0, [S - 1], rewards[b])
        close(d[b], ref)


def test_32x32_against_sparse_oracle(variant):
    """C4-sized world (1 024 states), goal-directed reward, causal path."""
    n = 32
    S = n * n
    mdp = SP.icy_gridworld_sparse(n, 0.2)
    r = np.full(S, -0.1); r[S - 1] = 1.0
    p0 = np.zeros(S); p0[0] = 1.0
    pa, n_lap = SP.local_causal_action_probabilities(mdp, [S - 1], r, 0.9)
    dref, n_svf = SP.expected_svf_from_policy(mdp, p0, [S - 1], pa)
    t = E.gridworld_tables(n, 0.2)
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    pol = E.soft_vi(t, phi, r, 0.9, mode=variant)
    assert counts()[0] == n_lap
    close(pol[0], pa)
    d = E.svf(t, p0, mask, pol[0], mode=variant)
    assert counts()[0] == n_svf
    close(d[0], dref)
    rr = -np.log(4.0) + 0.01 * np.random.default_rng(0).standard_normal(S)
    pb = SP.local_action_probabilities(mdp, [S - 1], rr)
    close(E.backward(t, mask, rr, mode=variant)[0], pb)


def test_grid_mode_64x64_properties():
    """Size-independent properties at a size the dense reference cannot reach quickly:
    policy rows sum to 1; SVF satisfies its own fixed-point equation to eps;
    sum of visit counts is finite and >= 1."""
    n = 64
    S = n * n
    t = E.gridworld_tables(n, 0.2)
    r = np.full(S, -0.1); r[S - 1] = 1.0
    p0 = np.zeros(S); p0[0] = 1.0
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    pol = E.soft_vi(t, phi, r, 0.9, mode=E.MODE_GRID)
    n_grid = counts()[0]
    pol_c = E.soft_vi(t, phi, r, 0.9, mode=E.MODE_CTA)
    assert counts()[0] == n_grid
    assert (pol == pol_c).all(), "grid and CTA variants must agree bitwise"
    rows = pol[0].sum(dim=1).cpu().numpy()
    np.testing.assert_allclose(np.delete(rows, S - 1), 1.0, rtol=1e-9)
    d = E.svf(t, p0, mask, pol[0], mode=E.MODE_GRID)
    n_grid = counts()[0]
    d_c = E.svf(t, p0, mask, pol[0], mode=E.MODE_CTA)
    assert counts()[0] == n_grid and (d == d_c).all()
    mdp = SP.icy_gridworld_sparse(n, 0.2)
    keep = np.ones(S); keep[S - 1] = 0
    dn, poln = d[0].cpu().numpy(), pol[0].cpu().numpy()
    nxt = p0 + sum(mdp.transposed()[a].dot(keep * poln[:, a] * dn) for a in range(4))
    assert np.max(np.abs(nxt - dn)) <= 1e-5
    assert 1.0 <= dn.sum() < 1e6


# ------------------------------------------------------------------ slab mode ---

def test_slab_single_rank_matches_engine_and_oracle():
    """The CUDA slab backend (one rank, no process group) against the sparse oracle and the
    single-GPU engine: same sweep counts, 1e-10."""
    import slab
    n = 16
    S = n * n
    g = slab.SlabGrid(n, 0.2, icy=True, chunk=20)
    r = np.full(S, -0.1); r[S - 1] = 1.0
    phi = np.full(S, -np.inf); phi[S - 1] = 0.0
    p0 = np.zeros(S); p0[0] = 1.0
    mdp = SP.icy_gridworld_sparse(n, 0.2)
    pa, n_lap = SP.local_causal_action_probabilities(mdp, [S - 1], r, 0.9)
    dref, n_svf = SP.expected_svf_from_policy(mdp, p0, [S - 1], pa)
    pol, v = g.soft_vi(r, phi, 0.9, 1e-5)
    assert g.last_n_iter == n_lap and g.last_status == 0
    close(pol, pa)
    d = g.svf(p0, [S - 1], pol, 1e-5)
    assert g.last_n_iter == n_svf
    close(d, dref)
    t = E.gridworld_tables(n, 0.2)
    pol_e = E.soft_vi(t, E.terminal_phi([S - 1], S), r, 0.9)
    assert (pol_e[0] == pol).all(), "slab sweep and persistent kernel must agree bitwise"
    vref, n_vi = SP.value_iteration(mdp, r, 0.95, 1e-5)
    val = g.value_iteration(r, 0.95, 1e-5)
    assert g.last_n_iter == n_vi
    close(val, vref)


@pytest.mark.parametrize("flow", [True, False])
def test_peer_slab_single_rank(flow):
    """The persistent slab kernels (flow: neighbour flags + chunked stop rule; else: barrier per sweep)
    with one rank: same results and sweep counts as the oracle and bitwise the same as the single-GPU engine."""
    import slab
    n = 16
    S = n * n
    g = slab.PeerSlabGrid(n, 0.2, icy=True, flow=flow)
    try:
        r = np.full(S, -0.1); r[S - 1] = 1.0
        phi = np.full(S, -np.inf); phi[S - 1] = 0.0
        p0 = np.zeros(S); p0[0] = 1.0
        mdp = SP.icy_gridworld_sparse(n, 0.2)
        pa, n_lap = SP.local_causal_action_probabilities(mdp, [S - 1], r, 0.9)
        dref, n_svf = SP.expected_svf_from_policy(mdp, p0, [S - 1], pa)
        pol, v = g.soft_vi(r, phi, 0.9, 1e-5)
        assert g.last_n_iter == n_lap and g.last_status == 0
        close(pol, pa)
        d = g.svf(p0, [S - 1], pol, 1e-5)
        assert g.last_n_iter == n_svf
        close(d, dref)
        t = E.gridworld_tables(n, 0.2)
        pol_e = E.soft_vi(t, E.terminal_phi([S - 1], S), r, 0.9)
        assert (pol_e[0] == pol).all()
        vref, n_vi = SP.value_iteration(mdp, r, 0.95, 1e-5)
        val = g.value_iteration(r, 0.95, 1e-5)
        assert g.last_n_iter == n_vi
        close(val, vref)
        d = g.svf(p0, [S - 1], pol, 1e-5, max_sweeps=25)
        assert g.last_n_iter == 25 and g.last_status == E.ST_MAXSWEEPS
    finally:
        g.close()


@pytest.mark.parametrize("n,chunk,blocks", [(16, 1, 0), (16, 7, 0), (64, 32, 0), (64, 64, 3), (64, 5, 16),
                                            (200, 32, 0), (200, 13, 37), (96, 64, 1)])
def test_slab_flow_kernel_is_bitwise_the_barrier_kernel(n, chunk, blocks, monkeypatch):
    """Dataflow slab kernel (csrc/slab_flow.cu: neighbour progress flags, stop rule all-reduced once per
    chunk, snapshot + replay) against the barrier-per-sweep slab kernel and the single-GPU cooperative
    grid kernel: identical sweep counts and bitwise identical soft-VI policy / value, VI value and
    forward SVF for every chunk length and CTA count (stop inside a chunk, on its last sweep, guard hit
    inside / on the edge of a chunk)."""
    import slab
    if blocks:
        monkeypatch.setenv("IRLB200_FLOW_BLOCKS", str(blocks))
    S = n * n
    rng = np.random.default_rng(n + chunk)
    r = -0.1 + 0.02 * rng.standard_normal(S); r[S - 1] = 1.0
    phi = np.full(S, -np.inf); phi[S - 1] = 0.0
    p0 = np.zeros(S); p0[0] = 0.7; p0[S // 2] = 0.3
    gf = slab.PeerSlabGrid(n, 0.2, icy=True, flow=True, chunk_sweeps=chunk)
    gb = slab.PeerSlabGrid(n, 0.2, icy=True, flow=False)
    try:
        pol_f, v_f = gf.soft_vi(r, phi, 0.9, 1e-5)
        n_f = gf.last_n_iter
        pol_b, v_b = gb.soft_vi(r, phi, 0.9, 1e-5)
        assert n_f == gb.last_n_iter and gf.last_status == 0
        assert (pol_f == pol_b).all() and (v_f == v_b).all()
        t = E.gridworld_tables(n, 0.2, slots=4)
        pol_e = E.soft_vi(t, E.terminal_phi([S - 1], S), r, 0.9, mode=E.MODE_GRID)
        assert counts()[0] == n_f and (pol_e[0] == pol_f).all()
        # forward pass: to convergence at the small sizes, otherwise budgets around the chunk edges
        budgets = [None] if n <= 16 else [1, chunk, chunk + 1, 3 * chunk - 1, 257]
        for ms in budgets:
            d_f = gf.svf(p0, [S - 1], pol_f, 1e-5, max_sweeps=ms)
            nf, sf = gf.last_n_iter, gf.last_status
            d_b = gb.svf(p0, [S - 1], pol_f, 1e-5, max_sweeps=ms)
            assert (nf, sf) == (gb.last_n_iter, gb.last_status), (ms, nf, sf, gb.last_n_iter, gb.last_status)
            assert (d_f == d_b).all(), ms
        val_f = gf.value_iteration(r, 0.95, 1e-4)
        nv = gf.last_n_iter
        val_b = gb.value_iteration(r, 0.95, 1e-4)
        assert nv == gb.last_n_iter and (val_f == val_b).all()
        # NaN ends the loop at the sweep in which it first shows (status NONFINITE), like the reference
        rn = r.copy(); rn[S // 3] = np.nan
        val_f = gf.value_iteration(rn, 0.95, 1e-4)
        nf, sf = gf.last_n_iter, gf.last_status
        val_b = gb.value_iteration(rn, 0.95, 1e-4)
        assert (nf, sf) == (gb.last_n_iter, gb.last_status) and sf == E.ST_NONFINITE
    finally:
        gf.close()
        gb.close()


def test_launch_order_hint_does_not_change_results():
    """Longest-first launch order of a batch (irlb200_svf_ordered): a scheduling hint only."""
    import torch
    n, B = 16, 300
    S = n * n
    tabs = E.gridworld_tables(n, 0.1 + 0.2 * np.arange(B) / B)
    rng = np.random.default_rng(3)
    r = -np.log(4.0) + 0.01 * rng.standard_normal((B, S))
    p0 = np.zeros(S); p0[0] = 1.0
    mask = E.terminal_mask([S - 1], S)
    pol = E.backward(tabs, mask, r)
    d0 = E.svf(tabs, p0, mask, pol, 1e-5, order=None)
    n0 = counts().copy()
    assert getattr(tabs, "_svf_order", None) is None
    perm = torch.randperm(B, device=d0.device).to(torch.int32)
    d1, g1 = E.svf(tabs, p0, mask, pol, 1e-5, e_features=np.zeros(S), order=perm)
    assert (d1 == d0).all() and (counts() == n0).all() and (g1 == -d0).all()
    d2 = E.svf(tabs, p0, mask, pol, 1e-5)                       # "auto": records the order for the next call
    hint = tabs._svf_order.cpu().numpy()
    assert sorted(hint.tolist()) == list(range(B)) and (np.diff(n0[hint]) <= 0).all()
    d3 = E.svf(tabs, p0, mask, pol, 1e-5)                       # uses it
    assert (d2 == d0).all() and (d3 == d0).all() and (counts() == n0).all()


@pytest.mark.parametrize("n,streamed", [(48, "0"), (48, "1"), (200, "0"), (200, "1")])
def test_compact_four_slot_tables_give_bitwise_the_same_results(n, streamed, monkeypatch):
    """irlb200_gridworld_tables_k(K = 4): no entry lost (table contents equal the 5-slot tables' non-zero
    entries), and soft-VI / VI / forward pass / backward pass on them are bitwise those of the 5-slot
    tables in the cooperative-grid kernels (register-resident and streamed) -- the kernels that stream
    large worlds from HBM."""
    monkeypatch.setenv("IRLB200_FORCE_STREAMED", streamed)
    S = n * n
    t5, t4 = E.gridworld_tables(n, 0.2), E.gridworld_tables(n, 0.2, slots=4)
    assert (t4.Ks, t4.Kp, t4.stencil_n) == (4, 4, 0)
    for i5, p5, i4, p4 in ((t5.succ_idx, t5.succ_p, t4.succ_idx, t4.succ_p), (t5.pred_idx, t5.pred_p, t4.pred_idx, t4.pred_p)):
        i5, p5, i4, p4 = (x[0].cpu().numpy() for x in (i5, p5, i4, p4))
        assert (p5[:, 4, :] == 0).all()                          # the fifth slot is always padding
        assert (i5[:4] == i4).all() and (p5[:, :4, :] == p4).all()
    r = np.full(S, -0.1); r[S - 1] = 1.0
    p0 = np.zeros(S); p0[0] = 1.0
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    out = []
    for t in (t5, t4):
        pol = E.soft_vi(t, phi, r, 0.9, mode=E.MODE_GRID)
        n_lap = counts()[0]
        d = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=3000, mode=E.MODE_GRID)
        n_svf = counts()[0]
        v = E.value_iteration(t, r, 0.9, 1e-4, mode=E.MODE_GRID)
        n_vi = counts()[0]
        pb = E.backward(t, mask, np.full(S, -np.log(4.0)), n_sweeps=200, mode=E.MODE_GRID)
        out.append((pol, d, v, pb, n_lap, n_svf, n_vi))
    for a, b in zip(out[0][:4], out[1][:4]):
        # 200 backward sweeps do not reach every state of the 200 x 200 world: 0 / 0 there, in both
        assert np.array_equal(a.cpu().numpy(), b.cpu().numpy(), equal_nan=True)
    assert out[0][4:] == out[1][4:]
    # the batched one-CTA kernels take the compact tables too (generic gather)
    if n == 48:
        d1 = E.svf(t4, p0, mask, out[1][0], 1e-5, max_sweeps=3000, mode=E.MODE_CTA)
        assert (d1 == out[1][1]).all()


@pytest.fixture(params=["push", "barrier"])
def cluster_variant(request, monkeypatch):
    """push: st.async + mbarrier exchange (default); barrier: barrier.cluster + DSMEM loads."""
    monkeypatch.setenv("IRLB200_CLUSTER_PUSH", "1" if request.param == "push" else "0")
    return request.param


@pytest.mark.parametrize("n", [16, 64, 128])
def test_cluster_mode_forward_pass(n, cluster_variant):
    """Thread-block-cluster forward pass (both exchange variants) == cooperative-grid forward pass,
    bitwise, same count."""
    S = n * n
    t = E.gridworld_tables(n, 0.2)
    r = np.full(S, -0.1); r[S - 1] = 1.0
    p0 = np.zeros(S); p0[0] = 1.0
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    pol = E.soft_vi(t, phi, r, 0.9)
    budget = None if n <= 64 else 30000
    d_c = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=budget, mode=E.MODE_CLUSTER)
    n_c, st_c = counts()[0], E.last_info.stati()[0]
    d_g = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=budget, mode=E.MODE_GRID)
    assert counts()[0] == n_c and E.last_info.stati()[0] == st_c
    assert (d_c == d_g).all()
    if n == 16:
        mdp = SP.icy_gridworld_sparse(n, 0.2)
        dref, n_ref = SP.expected_svf_from_policy(mdp, p0, [S - 1], pol[0].cpu().numpy())
        assert n_c == n_ref
        close(d_c[0], dref)
    # batch of clusters
    d_b = E.svf(t, p0, mask, torch_stack2(pol), 1e-5, max_sweeps=budget, mode=E.MODE_CLUSTER)
    assert (d_b[0] == d_c[0]).all() and (d_b[1] == d_c[0]).all()


def torch_stack2(pol):
    import torch
    return torch.cat([pol, pol], 0)


@pytest.mark.parametrize("n,size", [(16, 2), (16, 4), (16, 8), (32, 16), (64, 2), (64, 16), (128, 16), (24, 4), (40, 2)])
def test_cluster_push_every_cluster_size(n, size, monkeypatch):
    """The push kernel at every cluster size (one tile row per CTA up to many; partly idle warps at
    24 x 24 / 40 x 40) against the cooperative-grid kernel: bitwise, same count, also when the
    max-sweeps guard stops it and when it stops on the first sweeps (huge eps)."""
    monkeypatch.setenv("IRLB200_CLUSTER_PUSH", "1")
    monkeypatch.setenv("IRLB200_CLUSTER_SIZE", str(size))
    S = n * n
    t = E.gridworld_tables(n, 0.2)
    r = -np.log(4.0) + 0.05 * np.random.default_rng(n).standard_normal(S)
    p0 = np.random.default_rng(n + 1).random(S); p0 /= p0.sum()
    mask = E.terminal_mask([S - 1, S // 2], S)
    pol = E.backward(t, mask, r)
    cases = [(1e-5, 700), (1e-5, 1), (1e-5, 2), (1e-5, 7), (1e-5, 8), (1e-5, 9), (1e-5, 16), (1e-5, 17), (0.5, None)]
    cases += [(c / S, 4000) for c in (0.5, 0.2, 0.05, 0.02)]           # runs of assorted lengths (guarded: slow mixing)
    for eps, budget in cases:
        d_c = E.svf(t, p0, mask, pol, eps, max_sweeps=budget, mode=E.MODE_CLUSTER)
        n_c, st_c = counts()[0], E.last_info.stati()[0]
        d_g = E.svf(t, p0, mask, pol, eps, max_sweeps=budget, mode=E.MODE_GRID)
        assert counts()[0] == n_c and E.last_info.stati()[0] == st_c, (eps, budget)
        assert (d_c == d_g).all(), (eps, budget)


@pytest.mark.parametrize("n,size", [(16, 2), (16, 8), (32, 4), (32, 16), (64, 8), (64, 2), (24, 4)])
@pytest.mark.parametrize("regime", ["neutral", "overflow", "underflow"])
def test_cluster_push_backward_equals_one_cta_tiled_kernel(n, size, regime, monkeypatch):
    """Backward pass spread over a cluster (st.async row exchange, cluster-wide rescale maxima) ==
    the one-CTA tiled kernel, bitwise, in every rescale regime and for the short sweep counts."""
    monkeypatch.setenv("IRLB200_CLUSTER_SIZE", str(size))
    S = n * n
    t = E.gridworld_tables(n, 0.2)
    rng = np.random.default_rng(7 * n + size)
    shift = {"neutral": -np.log(4.0), "overflow": 1.5, "underflow": -9.0}[regime]
    r = np.stack([shift + 0.05 * rng.standard_normal(S) for _ in range(2)])
    mask = E.terminal_mask([S - 1, S // 3], S)
    for n_sweeps in (None, 1, 2, 3, 64, 65):
        p_c = E.backward(t, mask, r, n_sweeps=n_sweeps, mode=E.MODE_CLUSTER)
        p_1 = E.backward(t, mask, r, n_sweeps=n_sweeps, mode=E.MODE_CTA)
        assert p_c.shape == p_1.shape == (2, S, 4)
        assert (p_c == p_1).all() or n_sweeps is not None and not np.isfinite(p_1.cpu().numpy()).all() and \
            (np.isnan(p_c.cpu().numpy()) == np.isnan(p_1.cpu().numpy())).all(), (regime, n_sweeps)
    assert np.isfinite(p_c.cpu().numpy()).all()


def test_cluster_push_backward_128(monkeypatch):
    """128 x 128 (BASELINE configs[2]): AUTO takes the cluster kernel; cluster policy == the per-action
    cooperative-grid sweeps to 1e-10 (a cross-variant check; the comparison with the oracle's range-extended
    loop at this size is test_c3_full_size_values_against_the_c_oracle)."""
    n = 128; S = n * n
    t = E.gridworld_tables(n, 0.2)
    r = -np.log(4.0) + 0.01 * np.random.default_rng(0).standard_normal(S)
    mask = E.terminal_mask([S - 1], S)
    p_auto = E.backward(t, mask, r)
    p_cl = E.backward(t, mask, r, mode=E.MODE_CLUSTER)
    assert (p_auto == p_cl).all()
    p_g = E.backward(t, mask, r, mode=E.MODE_GRID)
    close(p_cl, p_g.cpu().numpy())
    row = p_cl[0].cpu().numpy().sum(axis=1)
    np.testing.assert_allclose(row, 1.0, rtol=1e-12)


def test_cluster_push_nonfinite_stops(monkeypatch):
    """A NaN in p_initial ends the loop with status NONFINITE in the push kernel as in the others."""
    monkeypatch.setenv("IRLB200_CLUSTER_PUSH", "1")
    n = 64; S = n * n
    t = E.gridworld_tables(n, 0.2)
    mask = E.terminal_mask([S - 1], S)
    pol = E.backward(t, mask, np.full(S, -np.log(4.0)))
    p0 = np.zeros(S); p0[0] = 1.0; p0[S // 3] = np.nan
    E.svf(t, p0, mask, pol, 1e-5, max_sweeps=5000, mode=E.MODE_CLUSTER)
    assert E.last_info.stati()[0] == E.ST_NONFINITE and counts()[0] <= 64


# ------------------------------------------------------ randomized + edge cases ---

def _random_mdp(rng, S, A, K):
    P = np.zeros((S, S, A))
    for s in range(S):
        for a in range(A):
            succ = rng.choice(S, size=min(rng.integers(1, K + 1), S), replace=False)
            w = rng.random(len(succ)) + 0.05
            P[s, succ, a] = w / w.sum()
    return P


@pytest.mark.parametrize("seed", range(8))
def test_randomized_mdps_against_dense_oracle(seed, variant):
    """Random shapes (S, A, K), random rewards / start distributions / terminal sets: every kernel
    against the dense oracle (which is pinned bit-for-bit to the reference)."""
    rng = np.random.default_rng(1000 + seed)
    S, A, K = int(rng.integers(2, 70)), int(rng.integers(1, 7)), int(rng.integers(1, 6))
    P = _random_mdp(rng, S, A, K)
    term = sorted(set(int(v) for v in rng.choice(S, size=int(rng.integers(1, 3)), replace=False)))
    r = 0.4 * rng.standard_normal(S) - 0.3
    p0 = rng.random(S); p0 /= p0.sum()
    t = E.compress_dense(P)
    mask, phi = E.terminal_mask(term, S), E.terminal_phi(term, S)
    with np.errstate(invalid="ignore"):         # unreachable states are 0/0 in the reference too
        ref = D.local_action_probabilities(P, term, r, rescale=True)
    got = E.backward(t, mask, r, mode=variant)[0].cpu().numpy()
    ok = np.isfinite(ref)                       # states that cannot reach a terminal are 0/0 in both
    assert (np.isfinite(got) == ok).all()
    np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-10)
    g = 0.5 + 0.45 * rng.random()
    pc, n_ref = D.local_causal_action_probabilities(P, term, r, g, 1e-6)
    pol = E.soft_vi(t, phi, r, g, 1e-6, mode=variant)
    assert counts()[0] == n_ref
    close(pol[0], pc)
    damp = 0.9 * pc / np.maximum(pc.sum(axis=1, keepdims=True), 1e-300)
    dref, n_ref = D.expected_svf_from_policy(P, p0, term, damp, 1e-6)
    d = E.svf(t, p0, mask, damp, 1e-6, mode=variant)
    assert counts()[0] == n_ref
    close(d[0], dref)
    vref, n_ref = D.value_iteration(P, r, g, 1e-6)
    v = E.value_iteration(t, r, g, 1e-6, mode=variant)
    assert counts()[0] == n_ref
    close(v[0], vref)


def test_edge_cases():
    """One-state world, empty terminal set, zero start mass, huge eps, zero sweeps."""
    P1 = np.ones((1, 1, 2))
    t = E.compress_dense(P1)
    v = E.value_iteration(t, np.array([0.5]), 0.5, 1e-9)
    vref, n = D.value_iteration(P1, np.array([0.5]), 0.5, 1e-9)
    assert counts()[0] == n
    close(v[0], vref)
    P = D.icy_gridworld_table(4, 0.2)
    t = E.compress_dense(P)
    S = 16
    # zero start mass: the first sweep already changes nothing -> exactly one sweep
    d = E.svf(t, np.zeros(S), E.terminal_mask([15], S), np.full((S, 4), 0.25), 1e-5)
    assert counts()[0] == 1 and (d == 0).all()
    # huge eps: one sweep, result = p0
    p0 = np.zeros(S); p0[3] = 1.0
    d = E.svf(t, p0, E.terminal_mask([15], S), np.full((S, 4), 0.25), 10.0)
    dref, n = D.expected_svf_from_policy(P, p0, [15], np.full((S, 4), 0.25), 10.0)
    assert counts()[0] == n == 1
    close(d[0], dref)
    # empty terminal set: causal policy rows all sum to 1, value iteration unaffected
    pc, n = D.local_causal_action_probabilities(P, [], np.linspace(-1, 0, S), 0.8, 1e-6)
    pol = E.soft_vi(t, E.terminal_phi([], S), np.linspace(-1, 0, S), 0.8, 1e-6)
    assert counts()[0] == n
    close(pol[0], pc)
    # backward pass with an explicit sweep count (the reference hard-codes 2*S)
    ref = D.local_action_probabilities(P, [15], np.full(S, -1.0))
    close(E.backward(t, E.terminal_mask([15], S), np.full(S, -1.0), n_sweeps=2 * S)[0], ref)


def test_irl_on_a_large_lazy_world_with_implicit_features():
    """End-to-end irl / irl_causal on a world handle (no dense table, no dense features): trajectories
    from the sparse sampler, tables built on the device, identity features implicit.  Checked
    against the dense oracle at 12x12 with the same trajectories."""
    n = 12
    Sn = n * n
    world = W.IcyGridWorld(n, 0.2, dense=False)
    np.random.seed(3)
    r_true = np.full(Sn, -0.1); r_true[Sn - 1] = 1.0
    value = S.value_iteration(world.tables(), r_true, 0.9)
    policy = S.stochastic_policy_from_value(world, value, w=lambda x: np.exp(4 * x))
    initial = np.zeros(Sn); initial[0] = 1.0
    tjs = list(T.generate_trajectories(40, world, T.stochastic_policy_adapter(policy), initial, [Sn - 1]))
    assert world._dense is None, "the sampler must not build the dense table"
    F = W.state_features(world, implicit=True)
    opt = O.ExpSga(lr=O.linear_decay(lr0=0.2))
    r = M.irl_causal(world.tables(), F, [Sn - 1], tjs, opt, O.Constant(1.0), 0.9, eps=1e-3)
    # oracle on the dense table with the same statistics
    P = D.icy_gridworld_table(n, 0.2)
    ef = D.feature_expectation_from_trajectories(np.identity(Sn), tjs)
    p0 = D.initial_probabilities_from_trajectories(Sn, tjs)
    ref, n_ref, _ = D.irl_causal(P, np.identity(Sn), [Sn - 1], ef, p0, D.ExpSgaPort(D.linear_decay(0.2)), np.ones(Sn),
                                 0.9, eps=1e-3)
    assert opt.k == n_ref
    close(r, ref)


def test_irl_batch_equals_individual_runs(golden):
    """Hyper-parameter sweep in lockstep == the same candidates run one by one (same kernels):
    identical outer step counts, rewards equal to rounding; candidate 0 is main.py's setting and
    reproduces the reference's 375 steps."""
    g = golden("e2e_5x5")
    tabs = E.gridworld_tables(5, 0.2)
    ef, p0 = g["e_features"], g["p_initial"]
    make = [lambda: O.ExpSga(lr=O.linear_decay(lr0=0.2)), lambda: O.ExpSga(lr=O.linear_decay(lr0=0.1)),
            lambda: O.ExpSga(lr=O.power_decay(lr0=0.3)), lambda: O.ExpSga(lr=0.05),
            lambda: O.ExpSga(lr=O.exponential_decay(lr0=0.2, decay_rate=0.01))]
    rewards, steps = M.irl_batch(tabs, [24], ef, p0, [m() for m in make], O.Constant(1.0))
    assert steps[0] == int(g["irl_steps"]) == 375
    close(rewards[0], g["irl_reward"])
    for b, m in enumerate(make):
        r1, s1 = M.irl_batch(tabs, [24], ef, p0, [m()], O.Constant(1.0))
        assert s1[0] == steps[b]
        close(rewards[b], r1[0].cpu().numpy(), rtol=1e-12)
    # one elementwise optimizer stepping the whole batch at once
    efb = np.stack([ef, ef, ef])
    rv, sv = M.irl_batch(tabs, [24], efb, p0, O.ExpSga(lr=O.linear_decay(lr0=0.2)), O.Constant(1.0))
    assert list(sv) == [375, 375, 375]
    close(rv[2], g["irl_reward"])
    # causal variant with per-candidate worlds
    tabs2 = E.gridworld_tables(5, [0.2, 0.3])
    rc, sc = M.irl_batch(tabs2, [24], ef, p0, [make[0](), make[0]()], O.Constant(1.0), causal=True, discount=0.9)
    assert sc[0] == int(g["irl_causal_0.9_steps"])
    close(rc[0], g["irl_causal_0.9_reward"])


def test_error_paths_fail_loudly():
    """Bad requests raise EngineError with the library's message instead of computing something else."""
    t = E.gridworld_tables(6, [0.2, 0.3])
    S = 36
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    r = np.zeros((2, S))
    with pytest.raises(E.EngineError, match="one problem per call"):
        E.soft_vi(t, phi, r, 0.9, mode=E.MODE_GRID)                     # grid mode takes a single problem
    with pytest.raises(E.EngineError, match="cluster mode"):
        P = golden_random_P()
        E.svf(E.compress_dense(P), np.full(7, 1 / 7), E.terminal_mask([6], 7), np.full((7, 3), 0.3),
              mode=E.MODE_CLUSTER)                                      # no grid stencil -> no cluster kernel
    with pytest.raises(E.EngineError, match="batch"):
        E.backward(t, mask, np.zeros((3, S)))                           # 3 problems, 2 tables
    with pytest.raises(E.EngineError, match="trailing dimension"):
        E.backward(t, mask, np.zeros((2, S + 1)))
    with pytest.raises(E.EngineError, match="policy must have shape"):
        E.svf(t, np.zeros(S), mask, np.zeros((2, S, 3)))
    with pytest.raises(E.EngineError, match="p_transition must have shape"):
        E.compress_dense(np.zeros((4, 5, 2)))
    big_a = np.zeros((3, 3, 17)); big_a[:, 0, :] = 1.0
    with pytest.raises(E.EngineError, match="run-time A > 16"):
        E.value_iteration(E.compress_dense(big_a), np.zeros(3), 0.5)


def golden_random_P():
    P = np.zeros((7, 7, 3))
    for s in range(7):
        for a in range(3):
            P[s, (s + a + 1) % 7, a] = 0.6
            P[s, 6, a] += 0.4
    return P


# ------------------------------------------------------- full BASELINE sizes ---

def test_c3_full_size_counts_and_fixed_point():
    """C3, 128x128 (16 384 states), goal-directed reward, gamma = 0.9, eps 1e-5: the sweep counts the
    survey measured with the reference-validated restatement (1 208 soft-VI sweeps,
    243 869 forward sweeps), and the converged vector satisfies the reference's own update to eps."""
    n = 128
    Sn = n * n
    t = W.IcyGridWorld(n, 0.2).tables()
    r = np.full(Sn, -0.1); r[Sn - 1] = 1.0
    p0 = np.zeros(Sn); p0[0] = 1.0
    mask, phi = E.terminal_mask([Sn - 1], Sn), E.terminal_phi([Sn - 1], Sn)
    pol = E.soft_vi(t, phi, r, 0.9, 1e-5)
    assert counts()[0] == 1208
    rows = pol[0].sum(dim=1).cpu().numpy()
    np.testing.assert_allclose(np.delete(rows, Sn - 1), 1.0, rtol=1e-9)
    d = E.svf(t, p0, mask, pol, 1e-5)                      # AUTO -> thread-block-cluster kernel
    assert counts()[0] == 243869 and E.last_info.stati()[0] == E.ST_CONVERGED
    mdp = SP.icy_gridworld_sparse(n, 0.2)
    keep = np.ones(Sn); keep[Sn - 1] = 0.0
    dn, pn = d[0].cpu().numpy(), pol[0].cpu().numpy()
    nxt = p0 + sum(mdp.transposed()[a].dot(keep * pn[:, a] * dn) for a in range(4))
    assert np.max(np.abs(nxt - dn)) <= 1e-5
    # d[terminal] is the probability mass absorbed so far; the reference's loose stop rule
    # (delta <= 1e-5) ends the loop with part of the mass still under way at this size
    assert 0.5 < dn[Sn - 1] <= 1.0


@pytest.mark.timeout(900)
def test_c3_full_size_values_against_the_c_oracle(monkeypatch):
    """C3 at full size, VALUES against the plain-C oracle (oracle/c/irl_oracle.c) to 1e-10 with exact counts:
    soft-VI policy (cooperative grid AND thread-block cluster), forward SVF of the causal and of the MaxEnt
    body (cluster push kernel), and the cluster backward policy against the oracle's range-extended loop."""
    from oracle import c_port as C
    n = 128
    S = n * n
    t = E.gridworld_tables(n, 0.2)
    sidx, sp = C.ell_from_sparse(SP.icy_gridworld_sparse(n, 0.2))
    mask, phi_d = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    phi = np.full(S, -np.inf); phi[S - 1] = 0.0
    p0 = np.zeros(S); p0[0] = 1.0
    # causal body: goal-directed reward
    r = np.full(S, -0.1); r[S - 1] = 1.0
    pol_c, _, n_c = C.soft_vi(sidx, sp, phi, r, 0.9, 1e-5)
    for mode in (E.MODE_GRID, E.MODE_AUTO):
        pol = E.soft_vi(t, phi_d, r, 0.9, mode=mode)
        assert counts()[0] == n_c == 1208
        close(pol[0], pol_c)
    d_c, n_dc = C.svf(sidx, sp, p0, [S - 1], pol_c, 1e-5)
    d = E.svf(t, p0, mask, pol, 1e-5)                            # AUTO -> cluster push kernel
    assert counts()[0] == n_dc == 243869
    close(d[0], d_c)
    # MaxEnt body: reward near -ln 4 (the regime in which the raw reference is finite, SURVEY TL;DR 2)
    rm = -np.log(4.0) + 0.01 * np.random.default_rng(0).standard_normal(S)
    pb_c = C.backward(sidx, sp, [S - 1], rm, rescale=True)
    pb = E.backward(t, mask, rm)                                 # AUTO -> cluster push kernel
    close(pb[0], pb_c)
    monkeypatch.setenv("IRLB200_BWD_TILE", "0")                  # per-action sweeps behind the grid barrier
    pb_g = E.backward(t, mask, rm, mode=E.MODE_GRID)
    close(pb_g[0], pb_c)
    monkeypatch.delenv("IRLB200_BWD_TILE")
    dm_c, n_mc = C.svf(sidx, sp, p0, [S - 1], pb_c, 1e-5)
    dm = E.svf(t, p0, mask, pb, 1e-5)
    assert counts()[0] == n_mc
    close(dm[0], dm_c)


def test_c4_sampled_worlds_of_the_full_batch():
    """C4: the 4096-world batch of bench.py; three sampled worlds against the sparse oracle
    (values 1e-10, forward sweep counts exact)."""
    B, n = 4096, 32
    Sn = n * n
    ps = 0.1 + 0.2 * np.arange(B) / 4096.0
    tabs = E.gridworld_tables(n, ps)
    theta = -np.log(4.0) + 0.01 * np.random.default_rng(1000).standard_normal((B, Sn))
    p0 = np.zeros(Sn); p0[0] = 1.0
    d, _ = M.compute_expected_svf_batch(tabs, p0, [Sn - 1], theta, fused=False, max_sweeps=2000000)
    nb = E.last_info.counts()
    assert (E.last_info.stati()[:, 1] == E.ST_CONVERGED).all()
    for b in (0, 1777, 4095):
        mdp = SP.icy_gridworld_sparse(n, ps[b])
        pa = SP.local_action_probabilities(mdp, [Sn - 1], theta[b])
        dref, n_ref = SP.expected_svf_from_policy(mdp, p0, [Sn - 1], pa, 1e-5)
        assert nb[b, 1] == n_ref
        close(d[b], dref)
    # 32 more worlds, evenly spaced through the batch, against the plain-C restatement (all host cores)
    from oracle import c_port as C
    sample = np.arange(64, B, B // 32)[:32]
    tabs_c = [C.ell_from_sparse(SP.icy_gridworld_sparse(n, ps[b])) for b in sample]
    out, n_c = C.batch_maxent_step(np.stack([t[0] for t in tabs_c]), np.stack([t[1] for t in tabs_c]), [Sn - 1], p0,
                                   theta[sample])
    assert (nb[sample, 1] == n_c).all()
    close(d[sample], out)


def test_c5_full_size_sweeps_against_sparse_oracle():
    """C5, 2048x2048 (4.2 M states) on one GPU in streamed cooperative-grid mode: 40 sweeps of soft-VI
    and of the forward pass against the sparse oracle at full size (1e-10), plus mass bounds."""
    n = 2048
    Sn = n * n
    t = E.gridworld_tables(n, 0.2)
    rng = np.random.default_rng(5)
    r = -0.1 + 0.05 * rng.standard_normal(Sn); r[Sn - 1] = 1.0
    p0 = np.zeros(Sn); p0[0] = 0.5; p0[Sn // 2 + n // 2] = 0.5
    mask, phi = E.terminal_mask([Sn - 1], Sn), E.terminal_phi([Sn - 1], Sn)
    mdp = SP.icy_gridworld_sparse(n, 0.2)
    k = 40
    pol, val = E.soft_vi(t, phi, r, 0.9, max_sweeps=k, mode=E.MODE_GRID, want_value=True)
    assert counts()[0] == k and E.last_info.stati()[0] == E.ST_MAXSWEEPS
    _, _, vref = SP.local_causal_action_probabilities(mdp, [Sn - 1], r, 0.9, max_sweeps=k, return_value=True)
    # after 40 sweeps most values still carry -1e200 * 0.9^40, so exp(q - v) of a truncated run is not
    # a meaningful policy (in the reference neither): compare the iterate itself, relatively
    close(val[0], vref)
    uniform = np.full((Sn, 4), 0.25)
    d = E.svf(t, p0, mask, uniform, 1e-5, max_sweeps=k, mode=E.MODE_GRID)
    dref, _ = SP.expected_svf_from_policy(mdp, p0, [Sn - 1], uniform, 1e-5, max_sweeps=k)
    close(d[0], dref)
    dn = d[0].cpu().numpy()
    assert dn.min() >= 0.0 and dn.sum() <= k + 1e-9 and abs(dn.sum() - k) < 1e-6     # nothing absorbed yet


@pytest.mark.parametrize("n,k", [(8, 7), (8, 16), (12, 33), (32, 1)])
def test_tiled_kernels_respect_the_sweep_guard(n, k):
    """Stencil-tiled forward kernel stopped by the guard: exactly k sweeps, iterate == the oracle's
    k-sweep iterate; the merged-weight backward pass with an explicit sweep count == the oracle's."""
    Sn = n * n
    t = E.gridworld_tables(n, 0.25)
    assert t.stencil_n == n
    mdp = SP.icy_gridworld_sparse(n, 0.25)
    p0 = np.zeros(Sn); p0[0] = 0.7; p0[Sn // 2] = 0.3
    pol = np.random.default_rng(n).dirichlet(np.ones(4), size=Sn)
    mask = E.terminal_mask([Sn - 1], Sn)
    d = E.svf(t, p0, mask, pol, 1e-9, max_sweeps=k)
    assert counts()[0] == k and E.last_info.stati()[0] == E.ST_MAXSWEEPS
    dref, _ = SP.expected_svf_from_policy(mdp, p0, [Sn - 1], pol, 1e-9, max_sweeps=k)
    close(d[0], dref)
    # backward: k sweeps instead of 2S (dense oracle with the same count)
    P = D.icy_gridworld_table(n, 0.25)
    r = -1.2 + 0.1 * np.random.default_rng(k).standard_normal(Sn)
    er = np.exp(r)
    zs = np.zeros(Sn); zs[Sn - 1] = 1.0
    for _ in range(k):
        za = np.array([er * P[:, :, a].dot(zs) for a in range(4)]).T
        zs = za.sum(axis=1)
    with np.errstate(invalid="ignore"):
        ref = za / zs[:, None]
    got = E.backward(t, mask, r, n_sweeps=k)[0].cpu().numpy()
    ok = np.isfinite(ref)
    assert (np.isfinite(got) == ok).all()
    np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-10)


def test_compress_dense_at_c3_size():
    """Kernel (1) at the C3 size: the 8.59 GB dense table of IcyGridWorld(128) (built on the device)
    compressed to tables == the directly built tables, bit for bit.  K is discovered, not assumed: an
    icy grid state has at most 4 distinct successors / predecessors (4 neighbours, or 3 + itself on an
    edge), stored in the 5 stencil slots of the register-resident shape."""
    import torch
    n = 128
    Pd = E.gridworld_dense(n, 0.2)
    assert Pd.shape == (n * n, n * n, 4) and Pd.element_size() * Pd.numel() == 8589934592
    a = E.compress_dense(Pd)
    del Pd
    torch.cuda.empty_cache()
    b = E.gridworld_tables(n, 0.2)
    assert a.k_discovered == (4, 4) and (a.Ks, a.Kp) == (5, 5) and a.stencil_n == n
    for x, y in ((a.succ_idx, b.succ_idx), (a.succ_p, b.succ_p), (a.pred_idx, b.pred_idx), (a.pred_p, b.pred_p)):
        assert (x == y).all()


# ------------------------------------------------------------------ dense batch ---

def _dense_random_mdp(rng, S, A, absorbing):
    P = rng.random((S, S, A)) ** 3                      # dense, skewed: every state reaches every state
    P[absorbing, :, :] = 0.0
    P[absorbing, absorbing, :] = 1.0
    P /= P.sum(axis=1, keepdims=True)
    return P


@pytest.mark.parametrize("S,A,B", [(64, 4, 5), (200, 3, 70), (256, 4, 33)])
def test_dense_batch_path_against_the_dense_oracle(S, A, B):
    """BASELINE north_star (4): B candidates sharing a DENSE table as FP64 tensor-core contractions
    (csrc/dense_batch.cu).  Against oracle/dense_port.py (the reference's arithmetic) candidate by candidate:
    1e-10 on policies, values and SVF, identical sweep counts; and against the ELL kernels on the same table."""
    rng = np.random.default_rng(S + B)
    term = S - 1
    P = _dense_random_mdp(rng, S, A, term)
    dt = E.DenseTables(P)
    rewards = -0.2 + 0.1 * rng.standard_normal((B, S))
    rewards[:, term] = 0.5
    p0 = rng.random(S); p0 /= p0.sum()
    mask = E.terminal_mask([term], S)
    check = sorted(set([0, B // 2, B - 1]))
    # soft-VI
    pol, val = E.dense_soft_vi(dt, E.terminal_phi([term], S), rewards, 0.9, 1e-6)
    n_lap = counts().copy()
    for b in check:
        pref, k = D.local_causal_action_probabilities(P, [term], rewards[b], 0.9, 1e-6)
        assert n_lap[b] == k
        close(pol[b], pref)
    # forward pass of those policies
    d = E.dense_svf(dt, p0, mask, pol, 1e-7)
    n_fw = counts().copy()
    for b in check:
        dref, k = D.expected_svf_from_policy(P, p0, [term], pol[b].cpu().numpy(), 1e-7)
        assert n_fw[b] == k
        close(d[b], dref)
    # backward pass (rewards near -ln S keep the raw reference finite) and value iteration
    rb = -np.log(A) + 0.01 * rng.standard_normal((B, S))
    pb = E.dense_backward(dt, mask, rb)
    for b in check:
        close(pb[b], D.local_action_probabilities(P, [term], rb[b], rescale=True))
        np.testing.assert_allclose(pb[b].cpu().numpy().sum(axis=1), 1.0, rtol=1e-12)
    v = E.dense_value_iteration(dt, rewards, 0.8, 1e-6)
    n_vi = counts().copy()
    for b in check:
        vref, k = D.value_iteration(P, rewards[b], 0.8, 1e-6)
        assert n_vi[b] == k
        close(v[b], vref)
    # the ELL kernels (run-time K = S gather) on the same dense table: same counts, 1e-10
    t = E.compress_dense(P)
    pol_e = E.soft_vi(t, E.terminal_phi([term], S), rewards[check], 0.9, 1e-6)
    assert (counts() == n_lap[check]).all()
    close(pol_e, pol[check].cpu().numpy())
    # module API, with the fused gradient epilogue
    ef = rng.random(S)
    svf, grad = M.compute_expected_svf_dense_batch(P, p0, [term], rewards, eps=1e-7, causal=True, discount=0.9,
                                                   eps_lap=1e-6, e_features=ef)
    assert (svf == d).all() and (grad == E.to_device(ef) - svf).all()
    assert (counts()[:, 0] == n_lap).all() and (counts()[:, 1] == n_fw).all()


def test_batches_of_worlds_too_large_for_one_cta():
    """a batch of problems beyond one CTA's shared memory (the C3 size) used to fail with ELIMIT in
    soft-VI / VI; the wrappers now run it as single-problem launches with the same results and counts."""
    n = 128
    S = n * n
    tabs = E.gridworld_tables(n, [0.2, 0.3])
    r = np.stack([np.full(S, -0.1), np.full(S, -0.2)]); r[:, S - 1] = 1.0
    phi = E.terminal_phi([S - 1], S)
    pol = E.soft_vi(tabs, phi, r, 0.9)
    n_b = counts().copy()
    for b in range(2):
        one = E.soft_vi(tabs.select(b), phi, r[b], 0.9)
        assert counts()[0] == n_b[b] and (one[0] == pol[b]).all()
    v = E.value_iteration(tabs, r, 0.9, 1e-4)
    assert v.shape == (2, S) and (counts() > 10).all()
    d, g = M.compute_expected_svf_batch(tabs, np.eye(S)[0], [S - 1], r, causal=True, discount=0.9, max_sweeps=3000)
    assert d.shape == (2, S) and (counts()[:, 0] == n_b).all() and (counts()[:, 1] == 3000).all()
