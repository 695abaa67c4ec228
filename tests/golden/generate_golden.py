#!/usr/bin/env python
"""
Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/generate_golden.py

The reference modules are loaded from /root/reference/src under private module
names (so they cannot shadow the product modules of the same name), with the
single shim the survey documents: `np.float = float` (maxent.py:314,336 use the
alias numpy removed in 1.24).  Sweep counts are taken from the reference's own
loops by counting its calls to `np.max` through a forwarding proxy installed
as the module-level `np` of the loaded reference module -- the reference source
is not edited.

Nothing here is imported by the product path.
"""

import importlib.util
import os
import sys

import numpy as np

REF = os.environ.get("IRL_REFERENCE_SRC", "/root/reference/src")
OUT = os.path.dirname(os.path.abspath(__file__))

np.float = float        # documented shim


def _load(name):
    spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _CountingNumpy:
    """Forwards everything to numpy, counts calls of np.max."""

    def __init__(self):
        self.n_max = 0

    def __getattr__(self, item):
        return getattr(np, item)

    def max(self, *a, **k):
        self.n_max += 1
        return np.max(*a, **k)


W = _load("gridworld")
M = _load("maxent")
S = _load("solver")
O = _load("optimizer")
T = _load("trajectory")

_proxy_m = _CountingNumpy()
_proxy_s = _CountingNumpy()
M.np = _proxy_m
S.np = _proxy_s


def counted(fn, proxy, per_sweep, *a, **k):
    proxy.n_max = 0
    with np.errstate(all="ignore"):
        out = fn(*a, **k)
    assert proxy.n_max % per_sweep == 0
    return out, proxy.n_max // per_sweep


def svf(*a, **k):
    return counted(M.expected_svf_from_policy, _proxy_m, 1, *a, **k)


def lcap(*a, **k):
    return counted(M.local_causal_action_probabilities, _proxy_m, 1, *a, **k)


def vi(*a, **k):
    return counted(S.value_iteration, _proxy_s, 2, *a, **k)


def random_mdp(rng, n_states, n_actions, max_succ):
    """Random sparse MDP: every (s,a) row has 1..max_succ successors, rows sum to 1."""
    P = np.zeros((n_states, n_states, n_actions))
    for s in range(n_states):
        for a in range(n_actions):
            k = rng.integers(1, max_succ + 1)
            succ = rng.choice(n_states, size=min(k, n_states), replace=False)
            w = rng.random(len(succ)) + 0.05
            P[s, succ, a] = w / w.sum()
    return P


def gen_worlds():
    out = {}
    for n in (1, 2, 3, 4, 5, 6, 8):
        out["grid_%d" % n] = W.GridWorld(n).p_transition
        for p in (0.2, 0.35):
            out["icy_%d_%s" % (n, p)] = W.IcyGridWorld(n, p_slip=p).p_transition
    np.savez_compressed(os.path.join(OUT, "worlds.npz"), **out)
    print("worlds.npz", len(out))


def gen_kernels():
    """Per-function known-answer vectors (the SURVEY section 8c KATs and more)."""
    out = {}
    world = W.IcyGridWorld(5, 0.2)
    P = world.p_transition
    r = np.zeros(25); r[24] = 1.0; r[8] = 0.65
    p0 = np.zeros(25); p0[0] = 1.0
    out["k5_reward"], out["k5_p0"] = r, p0

    pa = M.local_action_probabilities(P, [24], r)
    out["k5_lap"] = pa
    d, n = svf(P, p0, [24], pa)
    out["k5_svf"], out["k5_svf_n"] = d, n
    for g in (0.7, 0.9):
        pc, n = lcap(P, [24], r, g)
        out["k5_lcap_%s" % g], out["k5_lcap_%s_n" % g] = pc, n
        d, n = svf(P, p0, [24], pc)
        out["k5_csvf_%s" % g], out["k5_csvf_%s_n" % g] = d, n
    v, n = vi(P, r, 0.7)
    out["k5_vi"], out["k5_vi_n"] = np.asarray(v), n
    v, n = vi(P, r, 0.9, 1e-6)
    out["k5_vi9"], out["k5_vi9_n"] = np.asarray(v), n

    # terminal reward given as an array (maxent.py:313-314)
    phi = np.full(25, -3.0); phi[24] = 0.5; phi[7] = -np.inf
    pc, n = lcap(P, phi, r, 0.8, 1e-6)
    out["k5_phi"], out["k5_lcap_phi"], out["k5_lcap_phi_n"] = phi, pc, n

    # two terminals, distributed start
    p0b = np.zeros(25); p0b[[0, 3, 11]] = [0.5, 0.25, 0.25]
    rb = np.linspace(-0.5, 0.4, 25)
    pa = M.local_action_probabilities(P, [24, 4], rb)
    d, n = svf(P, p0b, [24, 4], pa, 1e-7)
    out["k5b_reward"], out["k5b_p0"], out["k5b_lap"], out["k5b_svf"], out["k5b_svf_n"] = rb, p0b, pa, d, n

    # larger grids, reward near -ln 4 keeps the raw backward pass finite (SURVEY TL;DR 2)
    for nside in (8, 12):
        rng = np.random.default_rng(nside)
        world = W.IcyGridWorld(nside, 0.2)
        Sn = nside * nside
        rr = -np.log(4.0) + 0.01 * rng.standard_normal(Sn)
        p0n = np.zeros(Sn); p0n[0] = 1.0
        pa = M.local_action_probabilities(world.p_transition, [Sn - 1], rr)
        d, n = svf(world.p_transition, p0n, [Sn - 1], pa)
        pre = "k%d_" % nside
        out[pre + "reward"], out[pre + "lap"], out[pre + "svf"], out[pre + "svf_n"] = rr, pa, d, n
        rg = np.full(Sn, -0.1); rg[Sn - 1] = 1.0
        pc, n = lcap(world.p_transition, [Sn - 1], rg, 0.9)
        d, n2 = svf(world.p_transition, p0n, [Sn - 1], pc)
        out[pre + "greward"], out[pre + "lcap"], out[pre + "lcap_n"] = rg, pc, n
        out[pre + "csvf"], out[pre + "csvf_n"] = d, n2
        v, n = vi(world.p_transition, rg, 0.95, 1e-5)
        out[pre + "vi"], out[pre + "vi_n"] = np.asarray(v), n

    # overflow regime of the raw reference: reward == 1 at 13x13 -> NaN policy (SURVEY TL;DR 2)
    world = W.IcyGridWorld(13, 0.2)
    with np.errstate(all="ignore"):
        pa = M.local_action_probabilities(world.p_transition, [168], np.ones(169))
    out["k13_lap_overflow"] = pa
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **out)
    print("kernels.npz", len(out))


def gen_random():
    """Random non-grid MDPs: exercise K discovery, ragged rows, A != 4."""
    out = {}
    cases = [(3, 2, 2), (7, 3, 3), (16, 4, 4), (33, 5, 6), (40, 1, 3), (64, 4, 9)]
    for i, (Sn, A, K) in enumerate(cases):
        rng = np.random.default_rng(100 + i)
        P = random_mdp(rng, Sn, A, K)
        # make the last state absorbing-terminal reachable: every state leaks a bit into it
        term = [Sn - 1]
        r = 0.3 * rng.standard_normal(Sn) - 0.5
        p0 = rng.random(Sn); p0 /= p0.sum()
        pre = "r%d_" % i
        out[pre + "P"], out[pre + "reward"], out[pre + "p0"] = P, r, p0
        out[pre + "terminal"] = np.array(term)
        pa = M.local_action_probabilities(P, term, r)
        out[pre + "lap"] = pa
        pc, n = lcap(P, term, r, 0.85)
        out[pre + "lcap"], out[pre + "lcap_n"] = pc, n
        # forward pass needs an absorbing policy: use a discounted-looking policy
        # (rows of pc sum to < 1 only at terminal), guard with the causal policy
        pol = 0.9 * pc / np.maximum(pc.sum(axis=1, keepdims=True), 1e-300)
        d, n = svf(P, p0, term, pol)
        out[pre + "pol"], out[pre + "svf"], out[pre + "svf_n"] = pol, d, n
        v, n = vi(P, r, 0.9, 1e-6)
        out[pre + "vi"], out[pre + "vi_n"] = np.asarray(v), n
    np.savez_compressed(os.path.join(OUT, "random_mdps.npz"), **out)
    print("random_mdps.npz", len(out))


class _CountingOptim:
    def __init__(self, inner):
        self.inner, self.n = inner, 0

    def reset(self, p):
        self.inner.reset(p)

    def step(self, g, *a, **k):
        self.n += 1
        return self.inner.step(g, *a, **k)


def gen_e2e():
    """main.py's pipeline (main.py:14-93) with np.random.seed(0): KAT5."""
    out = {}
    np.random.seed(0)
    world = W.IcyGridWorld(size=5, p_slip=0.2)
    reward = np.zeros(world.n_states); reward[-1] = 1.0; reward[8] = 0.65
    terminal = [24]
    initial = np.zeros(world.n_states); initial[0] = 1.0
    value = S.value_iteration(world.p_transition, reward, 0.7)
    policy = S.stochastic_policy_from_value(world, value, w=lambda x: x ** 5)
    tjs = list(T.generate_trajectories(200, world, T.stochastic_policy_adapter(policy), initial, terminal))
    lens = np.array([len(t.transitions()) for t in tjs])
    flat = np.concatenate([np.array(t.transitions(), dtype=np.int64).reshape(-1, 3) for t in tjs])
    out["traj_len"], out["traj_flat"] = lens, flat
    out["expert_value"], out["expert_policy"] = np.asarray(value), policy
    features = W.state_features(world)
    out["e_features"] = M.feature_expectation_from_trajectories(features, tjs)
    out["p_initial"] = M.initial_probabilities_from_trajectories(world.n_states, tjs)

    opt = _CountingOptim(O.ExpSga(lr=O.linear_decay(lr0=0.2)))
    r = M.irl(world.p_transition, features, terminal, tjs, opt, O.Constant(1.0))
    out["irl_reward"], out["irl_steps"] = r, opt.n
    for g in (0.7, 0.9):
        opt = _CountingOptim(O.ExpSga(lr=O.linear_decay(lr0=0.2)))
        r = M.irl_causal(world.p_transition, features, terminal, tjs, opt, O.Constant(1.0), g)
        out["irl_causal_%s_reward" % g], out["irl_causal_%s_steps" % g] = r, opt.n
    # other optimizer families / feature matrices
    feats = W.coordinate_features(world)
    out["coord_features"] = feats
    opt = _CountingOptim(O.ExpSga(lr=O.linear_decay(lr0=0.1)))
    r = M.irl(world.p_transition, feats, terminal, tjs, opt, O.Constant(0.2), eps=1e-3)
    out["irl_expsga_coord_reward"], out["irl_expsga_coord_steps"] = r, opt.n
    # plain Sga drives the coordinate-feature reward up until the raw backward pass
    # overflows: the reference then returns all-NaN (NaN ends every loop)
    opt = _CountingOptim(O.Sga(lr=0.02))
    with np.errstate(all="ignore"):
        r = M.irl(world.p_transition, feats, terminal, tjs, opt, O.Constant(-0.3), eps=1e-3)
    out["irl_sga_coord_nan_reward"], out["irl_sga_coord_nan_steps"] = r, opt.n
    opt = _CountingOptim(O.ExpSga(lr=O.power_decay(lr0=0.3)).normalize_grad())
    r = M.irl_causal(world.p_transition, features, terminal, tjs, opt, O.Constant(1.0), 0.8, eps=1e-3)
    out["irl_causal_ng_reward"], out["irl_causal_ng_steps"] = r, opt.n
    opt = _CountingOptim(O.Sga(lr=O.exponential_decay(lr0=0.1, decay_rate=0.05)).normalize_grad())
    r = M.irl_causal(world.p_transition, features, terminal, tjs, opt, O.Constant(0.2), 0.8, eps=1e-3)
    out["irl_causal_sga_reward"], out["irl_causal_sga_steps"] = r, opt.n
    np.savez_compressed(os.path.join(OUT, "e2e_5x5.npz"), **out)
    print("e2e_5x5.npz", {k: int(v) for k, v in out.items() if k.endswith("_steps")})


def gen_optimizer():
    """optimizer.py schedules and steppers on a fixed gradient sequence."""
    out = {}
    rng = np.random.default_rng(7)
    grads = rng.standard_normal((6, 5))
    out["grads"] = grads
    ks = np.arange(0, 12)
    out["linear"] = np.array([O.linear_decay(0.2, 0.5, 3)(k) for k in ks])
    out["power"] = np.array([O.power_decay(0.2, 0.5, 2, 3)(k) for k in ks])
    out["expo"] = np.array([O.exponential_decay(0.2, 0.3, 2)(k) for k in ks])

    def run(opt, init):
        th = init(5)
        opt.reset(th)
        for g in grads:
            opt.step(g)
        return th
    out["sga"] = run(O.Sga(lr=0.1), O.Constant(0.5))
    out["sga_lin"] = run(O.Sga(lr=O.linear_decay(0.2)), O.Constant(0.5))
    out["expsga"] = run(O.ExpSga(lr=O.linear_decay(0.2)), O.Constant(1.0))
    out["expsga_norm"] = run(O.ExpSga(lr=0.1, normalize=True), O.Constant(1.0))
    out["ng_l2"] = run(O.Sga(lr=0.1).normalize_grad(), O.Constant(0.0))
    out["ng_l1"] = run(O.ExpSga(lr=O.exponential_decay(0.2)).normalize_grad(1), O.Constant(1.0))
    out["const_fn"] = O.Constant(lambda shape: 1.0 / shape)(4)
    np.savez_compressed(os.path.join(OUT, "optimizer.npz"), **out)
    print("optimizer.npz", len(out))


def gen_main_fixture():
    """The reference's driver, byte for byte (the caller that must run unchanged on the engine; run by
    tests/test_gpu_parity.py::test_the_reference_main_py_runs_unchanged_on_the_gpu), and its licence."""
    import hashlib
    import shutil
    shutil.copyfile(os.path.join(REF, "main.py"), os.path.join(OUT, "reference_main_py.fixture"))
    shutil.copyfile(os.path.join(os.path.dirname(REF), "LICENSE"), os.path.join(OUT, "reference_main_py.LICENSE"))
    print("reference_main_py.fixture sha256",
          hashlib.sha256(open(os.path.join(OUT, "reference_main_py.fixture"), "rb").read()).hexdigest())


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference not present at %s: fixtures can only be regenerated in the build container" % REF)
    gen_worlds()
    gen_kernels()
    gen_random()
    gen_optimizer()
    gen_e2e()
    gen_main_fixture()
