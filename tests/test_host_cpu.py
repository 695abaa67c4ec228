"""Host-side logic of the product modules, CPU only: world tables, optimizers
(numpy and torch-CPU tensors), trajectory statistics, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import gridworld as W
import optimizer as O
import trajectory as T
import _irlb200 as E

from test_oracle_golden import load_trajectories

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ worlds ---

@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 8])
def test_world_tables_match_reference(golden, n):
    g = golden("worlds")
    assert np.array_equal(W.GridWorld(n).p_transition, g["grid_%d" % n])
    for p in (0.2, 0.35):
        assert np.array_equal(W.IcyGridWorld(n, p_slip=p).p_transition, g["icy_%d_%s" % (n, p)])


def test_zero_structure_like_reference_test():
    """Same property the reference's only test checks (src/test_gridworld.py:11-54)."""
    for world in (W.GridWorld(5), W.IcyGridWorld(5)):
        n = world.size
        for f in range(world.n_states):
            for t in range(world.n_states):
                if abs(f % n - t % n) + abs(f // n - t // n) > 1:
                    assert (world.p_transition[f, t, :] == 0.0).all()


def test_world_helpers():
    w = W.IcyGridWorld(5)
    assert w.n_states == 25 and w.n_actions == 4 and w.actions == [(1, 0), (-1, 0), (0, 1), (0, -1)]
    assert w.state_index_to_point(7) == (2, 1) and w.state_point_to_index((2, 1)) == 7
    assert w.state_point_to_index_clipped((-1, 9)) == 20
    nxt = w.intended_next_states()
    for s in range(25):
        for a in range(4):
            assert nxt[s, a] == w.state_index_transition(s, a)
    assert np.array_equal(W.state_features(w), np.identity(25))
    cf = W.coordinate_features(w)
    assert cf.shape == (25, 5) and cf[7, 2] == 1 and cf[7, 1] == 1 and cf[6, 1] == 2
    assert repr(w) == "IcyGridWorld(size=5, p_slip=0.2)" and repr(W.GridWorld(3)) == "GridWorld(size=3)"


def test_large_world_is_lazy():
    w = W.IcyGridWorld(128)
    assert w._dense is None and w.n_states == 16384


# --------------------------------------------------------------- optimizers ---

def _run(opt, init, grads, as_torch):
    th = init(grads.shape[1])
    if as_torch:
        import torch
        th = torch.as_tensor(th)
        grads = torch.as_tensor(grads)
    opt.reset(th)
    for g in grads:
        opt.step(g)
    return th.numpy() if as_torch else th


@pytest.mark.parametrize("as_torch", [False, True])
def test_optimizers_match_reference(golden, as_torch):
    g = golden("optimizer")
    grads = g["grads"]
    tol = dict(rtol=1e-14, atol=0) if as_torch else dict(rtol=0, atol=0)
    cases = {
        "sga": (O.Sga(lr=0.1), O.Constant(0.5)),
        "sga_lin": (O.Sga(lr=O.linear_decay(0.2)), O.Constant(0.5)),
        "expsga": (O.ExpSga(lr=O.linear_decay(0.2)), O.Constant(1.0)),
        "expsga_norm": (O.ExpSga(lr=0.1, normalize=True), O.Constant(1.0)),
        "ng_l2": (O.Sga(lr=0.1).normalize_grad(), O.Constant(0.0)),
        "ng_l1": (O.ExpSga(lr=O.exponential_decay(0.2)).normalize_grad(1), O.Constant(1.0)),
    }
    for name, (opt, init) in cases.items():
        np.testing.assert_allclose(_run(opt, init, grads, as_torch), g[name], err_msg=name, **tol)


def test_schedules_and_initializers(golden):
    g = golden("optimizer")
    ks = np.arange(12)
    assert np.array_equal([O.linear_decay(0.2, 0.5, 3)(k) for k in ks], g["linear"])
    assert np.array_equal([O.power_decay(0.2, 0.5, 2, 3)(k) for k in ks], g["power"])
    assert np.array_equal([O.exponential_decay(0.2, 0.3, 2)(k) for k in ks], g["expo"])
    assert np.array_equal(O.Constant(lambda shape: 1.0 / shape)(4), g["const_fn"])
    np.random.seed(3)
    a = O.Uniform(0.1, 0.4)(6)
    np.random.seed(3)
    assert np.array_equal(a, np.random.uniform(size=6, low=0.1, high=0.4))


def test_optimizer_aliasing():
    """reset() must keep a reference, step() must update in place."""
    th = np.ones(3)
    opt = O.ExpSga(lr=0.5).normalize_grad()
    opt.reset(th)
    assert opt.parameters is th and opt.opt.parameters is th
    opt.step(np.array([1.0, 0.0, -1.0]))
    assert th[0] > 1.0 > th[2]


# ------------------------------------------------------------- trajectories ---

def test_trajectory_container():
    t = T.Trajectory([(0, 1, 5), (5, 2, 6)])
    assert list(t.states()) == [0, 5, 6] and t.transitions()[0][0] == 0
    assert repr(t) == "Trajectory([(0, 1, 5), (5, 2, 6)])"


def test_trajectory_generation_matches_reference_stream(golden):
    """Seeded generation consumes numpy's global RNG exactly like the reference (main.py:32-51)."""
    g = golden("e2e_5x5")
    world = W.IcyGridWorld(5, 0.2)
    np.random.seed(0)
    initial = np.zeros(25); initial[0] = 1.0
    tjs = list(T.generate_trajectories(200, world, T.stochastic_policy_adapter(g["expert_policy"]), initial, [24]))
    ref = load_trajectories(g)
    assert [t.transitions() for t in tjs] == [t.transitions() for t in ref]


def test_trajectory_statistics(golden):
    import maxent as M
    g = golden("e2e_5x5")
    tjs = load_trajectories(g)
    assert np.array_equal(M.feature_expectation_from_trajectories(np.identity(25), tjs), g["e_features"])
    assert np.array_equal(M.initial_probabilities_from_trajectories(25, tjs), g["p_initial"])


# ------------------------------------------------------------------- C ABI ---

def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "irl_maxent_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(irlb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert set(_declared_symbols()) == set(E.SIGNATURES)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (and as C++) on its own."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "irl_maxent_b200.h"\nint main(void) { irlb200_tables t; (void)t; return irlb200_version() > 0 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])


def test_library_loads_and_exports_every_symbol():
    """dlopen only -- no compute call (there is no GPU in the CPU test tier)."""
    if not os.path.exists(E.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("build_native", os.path.join(ROOT, "irl-maxent_b200", "build_native.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = ctypes.CDLL(E.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    E.load_library()
    assert E._lib.irlb200_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import maxent as M
    P = W.IcyGridWorld(3).p_transition
    with pytest.raises(E.EngineError):
        M.local_action_probabilities(P, [8], np.zeros(9))
    with pytest.raises(E.EngineError):
        E.gridworld_tables(4, 0.2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "irl-maxent_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), \
                    "%s mentions the oracle" % os.path.join(dirpath, f)


# ------------------------------------------ sparse sampler / implicit features ---

def test_sparse_rows_equal_dense_rows():
    for world in (W.IcyGridWorld(5, 0.2), W.IcyGridWorld(2, 0.35), W.GridWorld(4), W.IcyGridWorld(1, 0.2)):
        P = world.p_transition
        for s in range(world.n_states):
            for a in range(world.n_actions):
                idx, p = world.successors(s, a)
                nz = np.nonzero(P[s, :, a])[0]
                assert np.array_equal(idx, nz) and np.array_equal(p, P[s, nz, a])


def test_sparse_sampler_consumes_rng_like_dense_choice(golden):
    """The O(1) sampler reproduces the reference's seeded trajectories bit for bit."""
    g = golden("e2e_5x5")
    world = W.IcyGridWorld(5, 0.2)
    np.random.seed(0)
    initial = np.zeros(25); initial[0] = 1.0
    tjs = list(T.generate_trajectories(200, world, T.stochastic_policy_adapter(g["expert_policy"]), initial, [24]))
    assert [t.transitions() for t in tjs] == [t.transitions() for t in load_trajectories(g)]
    # and against numpy's own dense draw, state by state
    rng_state = np.random.get_state()
    a = T._choice_sparse(*world.successors(7, 2))
    np.random.set_state(rng_state)
    b = np.random.choice(range(25), p=world.p_transition[7, :, 2])
    assert a == b


def test_implicit_identity_features(golden):
    import maxent as M
    g = golden("e2e_5x5")
    tjs = load_trajectories(g)
    F = W.state_features(W.IcyGridWorld(5), implicit=True)
    assert F.shape == (25, 25) and M._is_identity(F)
    assert np.array_equal(M.feature_expectation_from_trajectories(F, tjs), g["e_features"])
    big = W.IcyGridWorld(128)
    assert isinstance(W.state_features(big), W.IdentityFeatures) and big._dense is None


def test_public_signatures_match_the_reference():
    """Drop-in contract: names, argument order, keyword names and defaults of
    the reference's public functions (note `eps_esvf` in irl vs `eps_svf` in irl_causal)."""
    import inspect
    import maxent as M
    import solver as S
    expected = {
        (M, "irl"): "(p_transition, features, terminal, trajectories, optim, init, eps=0.0001, eps_esvf=1e-05)",
        (M, "irl_causal"): "(p_transition, features, terminal, trajectories, optim, init, discount, eps=0.0001, "
                           "eps_svf=1e-05, eps_lap=1e-05)",
        (M, "compute_expected_svf"): "(p_transition, p_initial, terminal, reward, eps=1e-05)",
        (M, "compute_expected_causal_svf"): "(p_transition, p_initial, terminal, reward, discount, eps_lap=1e-05, "
                                            "eps_svf=1e-05)",
        (M, "local_action_probabilities"): "(p_transition, terminal, reward)",
        (M, "local_causal_action_probabilities"): "(p_transition, terminal, reward, discount, eps=1e-05)",
        (M, "expected_svf_from_policy"): "(p_transition, p_initial, terminal, p_action, eps=1e-05)",
        (M, "softmax"): "(x1, x2)",
        (M, "feature_expectation_from_trajectories"): "(features, trajectories)",
        (M, "initial_probabilities_from_trajectories"): "(n_states, trajectories)",
        (S, "value_iteration"): "(p, reward, discount, eps=0.001)",
        (S, "stochastic_value_iteration"): "(p, reward, discount, eps=0.001)",
        (S, "optimal_policy_from_value"): "(world, value)",
        (S, "optimal_policy"): "(world, reward, discount, eps=0.001)",
        (T, "generate_trajectory"): "(world, policy, start, final)",
        (T, "generate_trajectories"): "(n, world, policy, start, final)",
        (O, "linear_decay"): "(lr0=0.2, decay_rate=1.0, decay_steps=1)",
        (O, "power_decay"): "(lr0=0.2, decay_rate=1.0, decay_steps=1, power=2)",
        (O, "exponential_decay"): "(lr0=0.2, decay_rate=0.5, decay_steps=1)",
    }
    for (mod, name), sig in expected.items():
        assert str(inspect.signature(getattr(mod, name))) == sig, name
    assert str(inspect.signature(S.stochastic_policy_from_value)).startswith("(world, value, w=")
    assert str(inspect.signature(O.ExpSga.__init__)) == "(self, lr, normalize=False)"
    assert str(inspect.signature(O.Sga.__init__)) == "(self, lr)"
    assert str(inspect.signature(O.Uniform.__init__)) == "(self, low=0.0, high=1.0)"
    assert str(inspect.signature(O.Constant.__init__)) == "(self, value=1.0)"
    assert str(inspect.signature(W.IcyGridWorld.__init__)).startswith("(self, size, p_slip=0.2")


def test_reference_main_resolves_to_the_engine():
    """The reference's unmodified main.py, launched through scripts/run_reference_main.py, imports the
    ENGINE's modules and runs until the first kernel call -- which, on a box without a GPU, must be
    the engine's loud failure (main.py:46 `S.value_iteration`).  Skipped where the reference is absent."""
    import subprocess
    import sys
    ref_main = "/root/reference/src/main.py"
    if not os.path.exists(ref_main):
        pytest.skip("reference not present on this box")
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run would simply succeed (covered by test_irl_5x5_like_main)")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_reference_main.py"), ref_main],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    assert "EngineError" in p.stderr and "no CPU fallback" in p.stderr
    assert "irl-maxent_b200/solver.py" in p.stderr and "value_iteration" in p.stderr


def test_terminal_must_be_indices_for_the_forward_pass():
    """A terminal-reward array is only valid for the causal policy pass (reference: maxent.py:99 breaks too)."""
    with pytest.raises(E.EngineError, match="state indices"):
        E.terminal_mask(np.full(9, -1.5), 9)
    with pytest.raises(E.EngineError, match="state indices"):
        E.terminal_mask([3, 9], 9)


def test_table_cache_fingerprint_sees_in_place_edits():
    """the dense-table cache must not hand out stale tables after an in-place edit.  The
    fingerprint covers every byte up to 64 MiB (also > 1 MiB, the old limit), a strided sample beyond,
    and torch tensors through their version counter."""
    import torch
    small = np.zeros((25, 25, 4)); small[0, 1, 2] = 0.5
    f0 = E._fingerprint(small)
    assert E._fingerprint(small) == f0
    small[3, 4, 1] = 0.25
    assert E._fingerprint(small) != f0
    mid = np.zeros((300, 300, 4))                       # 2.9 MB: was identity-only in round 1
    f0 = E._fingerprint(mid)
    mid[299, 17, 3] = 1e-9
    assert E._fingerprint(mid) != f0
    big = np.zeros((1500, 1500, 4))                     # 72 MB: sampled
    f0 = E._fingerprint(big)
    assert f0[0] == "s" and E._fingerprint(big) == f0
    big[:] = 0.125
    assert E._fingerprint(big) != f0
    t = torch.zeros(8, 8, 4, dtype=torch.float64)
    f0 = E._fingerprint(t)
    t[1, 2, 3] = 1.0
    assert E._fingerprint(t) != f0
    import maxent as M
    assert M.invalidate is E.invalidate and M.clear_cache is E.clear_cache


def test_device_loop_is_taken_only_for_untouched_builtin_optimizers():
    """The device-side outer loop (irlb200_irl_small) knows the update rule of plain `Sga` / `ExpSga` only: a
    wrapper, a subclass, `normalize=True`, NormalizeGrad or a patched `step` must keep the generic host loop."""
    import maxent as M
    assert M._builtin_optimizer_kind(O.Sga(lr=0.1)) == 0
    assert M._builtin_optimizer_kind(O.ExpSga(lr=O.linear_decay(0.2))) == 1
    assert M._builtin_optimizer_kind(O.ExpSga(lr=0.1, normalize=True)) is None
    assert M._builtin_optimizer_kind(O.Sga(lr=0.1).normalize_grad()) is None

    class Mine(O.ExpSga):
        pass

    class Wrapped:
        def __init__(self, inner):
            self.inner = inner

        def reset(self, p):
            self.inner.reset(p)

        def step(self, g):
            self.inner.step(g)

    assert M._builtin_optimizer_kind(Mine(lr=0.1)) is None
    assert M._builtin_optimizer_kind(Wrapped(O.ExpSga(lr=0.1))) is None
    orig = O.ExpSga.step
    try:
        O.ExpSga.step = lambda self, grad, *a, **k: orig(self, grad, *a, **k)
        assert M._builtin_optimizer_kind(O.ExpSga(lr=0.1)) is None
    finally:
        O.ExpSga.step = orig
    assert M._builtin_optimizer_kind(O.ExpSga(lr=0.1)) == 1
    # the learning rates handed to the kernel are the schedule's own values, step counter untouched
    opt = O.ExpSga(lr=O.linear_decay(lr0=0.2))
    opt.reset(np.ones(3))
    assert [O._rate(opt.lr, opt.k + i) for i in range(3)] == [0.2, 0.1, 0.2 / 3] and opt.k == 0
