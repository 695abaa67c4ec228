"""
maxent -- Maximum-Entropy (Ziebart 2008) and Maximum-Causal-Entropy (Ziebart
2010) inverse reinforcement learning on B200.

Drop-in for the reference module of the same name (`/root/reference/src/maxent.py`):
same function names, argument order, keyword names and defaults, same return
values.  The three fixed-point loops (partition-function sweeps :154-156, soft
value iteration :326-338, state-visitation sweeps :108-112) run as persistent
sm_100a kernels behind the C ABI in include/irl_maxent_b200.h; this module only
moves arguments to the device and back.

Type rule: numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.
`p_transition` may be the reference's dense `[S, S', A]` table (numpy or CUDA
tensor; compressed once and cached) or a pre-built `_irlb200.Tables` handle
(e.g. `gridworld.IcyGridWorld(...).tables()`), which is the only option for
state counts whose dense table cannot exist.  The cache notices in-place edits
of the table between calls (content fingerprint: every byte up to 64 MiB, a
4 M-element sample beyond; tensors by version counter); after editing a larger
array in place call `maxent.invalidate(p_transition)` or `maxent.clear_cache()`.

Differences from the reference, by design:
  * the non-causal backward pass is range-extended (exact power-of-two
    rescaling): where the raw loop overflows to NaN (n >= 13 with reward >= 0)
    this module returns the finite, correctly normalised policy; where the raw
    loop is finite the results agree to the last few ulp;
  * every convergence loop has a very large max-sweep guard
    (`_irlb200.DEFAULT_MAX_SWEEPS`) instead of looping forever;
  * there is no CPU fallback.
"""

import os

import numpy as np

import _irlb200 as E
from _irlb200 import clear_cache, invalidate      # noqa: F401  (table-cache control, see the module docstring)


# -- helpers -------------------------------------------------------------------

def _wants_tensor(*xs):
    return any(E.is_tensor(x) for x in xs)


def _out(t, as_tensor):
    return t if as_tensor else t.cpu().numpy()


def _is_identity(features):
    if type(features).__name__ == "IdentityFeatures":
        return True
    if E.is_tensor(features):
        return False        # checked on the host only; tensors go through the dense kernels
    f = np.asarray(features)
    return f.ndim == 2 and f.shape[0] == f.shape[1] and np.array_equal(f, np.identity(f.shape[0]))


# -- common functions (host side, once per irl call) -----------------------------

def feature_expectation_from_trajectories(features, trajectories):
    """Mean over trajectories of the summed feature rows of every visited state
    (reference: maxent.py:15-39).  Accumulates in visiting order, so the result is
    bit-identical with the reference's running sum.  Host side, once per irl call."""
    if type(trajectories).__name__ == "DeviceTrajectories":
        # visit counts were accumulated by the sampling kernel: sum_t sum_s features[s, :] = counts . features
        counts = trajectories.visit_counts.cpu().numpy()
        if _is_identity(features):
            return counts / len(trajectories)                   # integer counts: bit-identical to the running sum
        f = features.cpu().numpy() if E.is_tensor(features) else np.asarray(features)
        return counts.dot(f) / len(trajectories)
    if type(features).__name__ == "IdentityFeatures":
        # identity rows: the running sum of one-hot rows is a visit count (exact in float64)
        fe = np.zeros(features.shape[1])
        n = 0
        for t in trajectories:
            for s in t.states():
                fe[s] += 1.0
            n += 1
        return fe / n
    f = features.cpu().numpy() if E.is_tensor(features) else np.asarray(features)
    fe = np.zeros(f.shape[1])
    n = 0
    for t in trajectories:
        for s in t.states():
            fe += f[s, :]
        n += 1
    return fe / n


def initial_probabilities_from_trajectories(n_states, trajectories):
    """Empirical start-state distribution (reference: maxent.py:42-60)."""
    if type(trajectories).__name__ == "DeviceTrajectories":
        return trajectories.start_counts.cpu().numpy() / len(trajectories)
    p = np.zeros(n_states)
    n = 0
    for t in trajectories:
        p[t.transitions()[0][0]] += 1.0
        n += 1
    return p / n


# -- forward pass ----------------------------------------------------------------

def expected_svf_from_policy(p_transition, p_initial, terminal, p_action, eps=1e-5):
    """Expected state visitation frequencies of a policy (reference: maxent.py:63-114).
    `terminal` must be a collection of state indices (as in the reference, :99)."""
    tables = E.as_tables(p_transition)
    as_t = _wants_tensor(p_initial, p_action)
    mask = E.terminal_mask(terminal, tables.S)
    d = E.svf(tables, p_initial, mask, p_action, eps)
    return _out(d[0], as_t)


# -- plain maximum entropy ---------------------------------------------------------

def local_action_probabilities(p_transition, terminal, reward):
    """Backward pass of MaxEnt IRL, 2*S partition sweeps (reference: maxent.py:119-159)."""
    tables = E.as_tables(p_transition)
    mask = E.terminal_mask(terminal, tables.S)
    pol = E.backward(tables, mask, reward)
    return _out(pol[0], _wants_tensor(reward))


def compute_expected_svf(p_transition, p_initial, terminal, reward, eps=1e-5):
    """Backward pass + forward pass in one launch (reference: maxent.py:162-193)."""
    tables = E.as_tables(p_transition)
    mask = E.terminal_mask(terminal, tables.S)
    d, _, _ = E.expected_svf(tables, p_initial, mask, reward, causal=False, eps_svf=eps)
    return _out(d[0], _wants_tensor(p_initial, reward))


# -- maximum causal entropy ----------------------------------------------------------

def softmax(x1, x2):
    """Soft maximum max + log(1 + exp(min - max)) (reference: maxent.py:260-276).
    Elementwise helper kept for API compatibility; the kernels fold it on chip."""
    if _wants_tensor(x1, x2):
        torch = E._torch()
        x1, x2 = torch.as_tensor(x1), torch.as_tensor(x2)
        hi, lo = torch.maximum(x1, x2), torch.minimum(x1, x2)
        return hi + torch.log(1.0 + torch.exp(lo - hi))
    hi, lo = np.maximum(x1, x2), np.minimum(x1, x2)
    return hi + np.log(1.0 + np.exp(lo - hi))


def local_causal_action_probabilities(p_transition, terminal, reward, discount, eps=1e-5):
    """Soft value iteration + causal policy (reference: maxent.py:279-341).
    `terminal`: index collection, or the terminal reward function iff len == S."""
    tables = E.as_tables(p_transition)
    phi = E.terminal_phi(terminal, tables.S)
    pol = E.soft_vi(tables, phi, reward, discount, eps)
    return _out(pol[0], _wants_tensor(reward))


def compute_expected_causal_svf(p_transition, p_initial, terminal, reward, discount,
                                eps_lap=1e-5, eps_svf=1e-5):
    """Soft-VI + forward pass in one launch (reference: maxent.py:344-380)."""
    tables = E.as_tables(p_transition)
    mask = E.terminal_mask(terminal, tables.S)
    phi = E.terminal_phi(terminal, tables.S)
    d, _, _ = E.expected_svf(tables, p_initial, mask, reward, causal=True, phi=phi, discount=discount,
                             eps_lap=eps_lap, eps_svf=eps_svf)
    return _out(d[0], _wants_tensor(p_initial, reward))


# -- outer gradient loops ------------------------------------------------------------

def _builtin_optimizer_kind(optim):
    """0 / 1 if `optim` is this package's plain Sga / ExpSga (update rule known to irlb200_irl_small), with
    `step` and `reset` untouched -- a subclass, a wrapper or a patched class keeps the generic host loop."""
    import optimizer as O
    t = type(optim)
    if t is O.Sga and t.step is _SGA_STEP:
        return 0
    if t is O.ExpSga and t.step is _EXPSGA_STEP and not optim.normalize:
        return 1
    return None


def _capture_builtin_steps():
    import optimizer as O
    return O.Sga.step, O.ExpSga.step


_SGA_STEP, _EXPSGA_STEP = _capture_builtin_steps()
DEVICE_LOOP_CHUNK = 1024        # learning rates handed to one launch of the device-side outer loop


def _irl_device_loop(tables, mask, p_initial_d, e_features_d, theta, optim, kind, eps, causal, phi, discount,
                     eps_lap, eps_svf):
    """`while delta > eps` (maxent.py:240-252 / :436-450) inside irlb200_irl_small: the host evaluates the
    schedule for the next DEVICE_LOOP_CHUNK steps and reads back (steps, done) once per launch."""
    import optimizer as O
    torch = E._torch()
    th = theta.view(1, -1)
    steps = torch.zeros(1, dtype=torch.int32, device=theta.device)
    done = torch.zeros(1, dtype=torch.int32, device=theta.device)
    while True:
        rates = np.array([O._rate(optim.lr, optim.k + i) for i in range(DEVICE_LOOP_CHUNK)], dtype=np.float64)
        before = int(steps.item())
        counts = E.irl_small(tables, th, e_features_d, p_initial_d, mask, kind, rates, eps, steps, done, causal=causal,
                             phi=phi, discount=discount, eps_lap=eps_lap, eps_svf=eps_svf)
        optim.k += int(steps.item()) - before            # the optimizer object ends up where the host loop leaves it
        E.last_info = E.SweepInfo(counts, torch.zeros_like(counts))
        if int(done.item()):
            return


def _irl_loop(p_transition, features, terminal, trajectories, optim, init, eps, step, device_loop=None):
    """Shared body of irl / irl_causal (reference: maxent.py:229-255, :426-453).

    omega (`theta`) lives on the device for the whole optimisation; the optimizer
    steps it in place through the reference it was handed in `reset` (the
    reference relies on exactly this aliasing, :236-252).  The host blocks once
    per gradient step, on the scalar `delta`.
    """
    torch = E.require_cuda()
    tables = E.as_tables(p_transition)
    S = tables.S
    as_t = _wants_tensor(features)
    identity = _is_identity(features)
    n_features = features.shape[1]

    e_features = feature_expectation_from_trajectories(features, trajectories)
    p_initial = initial_probabilities_from_trajectories(S, trajectories)
    e_features_d, p_initial_d = E.to_device(e_features), E.to_device(p_initial)
    features_d = None if identity else E.to_device(features)
    mask = E.terminal_mask(terminal, S)

    theta = E.to_device(init(n_features))
    optim.reset(theta)
    kind = _builtin_optimizer_kind(optim)
    if (device_loop is not None and identity and kind is not None and S <= 32 and tables.A == 4
            and tables.Ks == 5 and tables.Kp == 5 and tables.n_tables == 1
            and os.environ.get("IRLB200_DEVICE_LOOP", "1") != "0"):
        # tiny world, identity features, built-in optimizer: the whole loop in one launch per 1024 steps
        _irl_device_loop(tables, mask, p_initial_d, e_features_d, theta, optim, kind, eps, **device_loop)
        return _out(theta.clone(), as_t)
    delta = np.inf
    while delta > eps:
        theta_old = theta.clone()
        # identity features: features.dot(theta) is theta itself, bit for bit
        reward = theta if identity else E.features_dot(features_d, theta)
        e_svf, grad = step(tables, p_initial_d, mask, reward, e_features_d if identity else None)
        if not identity:
            grad = E.features_grad(features_d, e_svf[0], e_features_d)
        else:
            grad = grad[0]
        optim.step(grad)
        # max |theta_old - theta| (NaN propagates, which ends the loop like the reference's np.max)
        delta = torch.linalg.vector_norm(theta_old - theta, ord=float("inf")).item()   # the one host sync per step

    reward = theta.clone() if identity else E.features_dot(features_d, theta)
    return _out(reward, as_t)


def irl(p_transition, features, terminal, trajectories, optim, init, eps=1e-4, eps_esvf=1e-5):
    """Maximum-entropy IRL; returns the per-state reward `features . theta`
    (reference: maxent.py:196-255)."""
    def step(tables, p0, mask, reward, ef):
        d, g, _ = E.expected_svf(tables, p0, mask, reward, causal=False, eps_svf=eps_esvf, e_features=ef)
        return d, g
    return _irl_loop(p_transition, features, terminal, trajectories, optim, init, eps, step,
                     device_loop=dict(causal=False, phi=None, discount=0.0, eps_lap=1e-5, eps_svf=eps_esvf))


def irl_causal(p_transition, features, terminal, trajectories, optim, init, discount,
               eps=1e-4, eps_svf=1e-5, eps_lap=1e-5):
    """Maximum-causal-entropy IRL; returns the per-state reward
    (reference: maxent.py:383-453)."""
    S = E.as_tables(p_transition).S
    phi = E.terminal_phi(terminal, S)
    # the forward pass needs an index list (reference: maxent.py:99); an array-valued
    # `terminal` is only meaningful for the policy pass, exactly as in the reference

    def step(tables, p0, mask, reward, ef):
        d, g, _ = E.expected_svf(tables, p0, mask, reward, causal=True, phi=phi, discount=discount,
                                 eps_lap=eps_lap, eps_svf=eps_svf, e_features=ef)
        return d, g
    return _irl_loop(p_transition, features, terminal, trajectories, optim, init, eps, step,
                     device_loop=dict(causal=True, phi=phi, discount=discount, eps_lap=eps_lap, eps_svf=eps_svf))


# -- batched mode (no counterpart in the reference: B independent problems) ------------

def compute_expected_svf_batch(tables, p_initial, terminal, reward, eps=1e-5, causal=False,
                               discount=None, eps_lap=1e-5, e_features=None, fused=None, max_sweeps=None):
    """B independent gradient-step bodies in one or two launches.

    tables: `Tables` with 1 (shared) or B worlds; reward [B,S]; p_initial [S] or [B,S];
    terminal: index collection shared by all problems.  Returns (svf [B,S], grad or None).
    """
    S = tables.S
    mask = E.terminal_mask(terminal, S)
    phi = E.terminal_phi(terminal, S) if causal else None
    d, g, _ = E.expected_svf(tables, p_initial, mask, reward, causal=causal, phi=phi,
                             discount=discount if causal else 0.0, eps_lap=eps_lap, eps_svf=eps,
                             e_features=e_features, fused=fused, max_sweeps=max_sweeps)
    return d, g


def compute_expected_svf_dense_batch(p_transition, p_initial, terminal, reward, eps=1e-5, causal=False,
                                     discount=None, eps_lap=1e-5, e_features=None, max_sweeps=None):
    """B reward candidates over ONE dense p_transition[S, S', A] (BASELINE configs[3], dense case): the
    gradient-step body of every candidate as FP64 tensor-core contractions [rows x S] . [S x B]
    (csrc/dense_batch.cu) -- for tables that really are dense (K ~ S); sparse worlds are faster through
    `compute_expected_svf_batch`.  `p_transition`: dense array / tensor or a prepared `_irlb200.DenseTables`.
    reward [B, S]; returns (svf [B, S], grad or None); per-candidate sweep counts in `_irlb200.last_info`."""
    torch = E.require_cuda()
    dt = p_transition if isinstance(p_transition, E.DenseTables) else E.DenseTables(p_transition)
    mask = E.terminal_mask(terminal, dt.S)
    if causal:
        pol, _ = E.dense_soft_vi(dt, E.terminal_phi(terminal, dt.S), reward, discount, eps_lap, max_sweeps)
        info_a = E.last_info
    else:
        pol = E.dense_backward(dt, mask, reward)
        info_a = None
    res = E.dense_svf(dt, p_initial, mask, pol, eps, max_sweeps, e_features)
    n_fw, st_fw = E.last_info.n_iter, E.last_info.status
    n_pol = info_a.n_iter if causal else torch.full_like(n_fw, 2 * dt.S)
    st_pol = info_a.status if causal else torch.zeros_like(st_fw)
    E.last_info = E.SweepInfo(torch.stack([n_pol, n_fw], 1), torch.stack([st_pol, st_fw], 1))
    return res if e_features is not None else (res, None)


def irl_batch(tables, terminal, e_features, p_initial, optims, init, eps=1e-4, eps_esvf=1e-5,
              causal=False, discount=None, eps_lap=1e-5, max_steps=None):
    """B independent IRL problems with identity features, run in lockstep on one GPU
    (hyper-parameter sweeps over optimizers / schedules / statistics).

    Candidate b has its own statistics `e_features[b]`, `p_initial[b]` ([B,S], or [S] shared) and the
    tables `tables` (one shared world, or B worlds).  `optims` is either
      * a list of B optimizers (anything with reset / step, e.g. `ExpSga` with different schedules):
        each steps ITS row of the device-resident omega in place, or
      * ONE optimizer whose step is elementwise (`Sga`, `ExpSga` without `normalize`, no
        `NormalizeGrad`): it steps the whole [B,S] omega at once; converged rows receive a zero
        gradient, which leaves them unchanged under both update rules.
    Every candidate follows exactly the loop of `irl` / `irl_causal` (maxent.py:240-252 / :436-450) and
    leaves the batch once max|omega_old - omega| <= eps.  With a shared world the still-active
    candidates are compacted before every launch.

    Returns (reward [B,S] device tensor, outer step counts [B] numpy).
    """
    torch = E.require_cuda()
    S = tables.S
    mask = E.terminal_mask(terminal, S)
    phi = E.terminal_phi(terminal, S) if causal else None
    ef = E.to_device(e_features)
    per_row = isinstance(optims, (list, tuple))
    B = len(optims) if per_row else (int(ef.shape[0]) if ef.dim() == 2 else tables.n_tables)
    ef = ef.expand(B, S).contiguous() if ef.dim() == 1 else ef
    p0 = E.to_device(p_initial)
    p0_shared = p0.dim() == 1
    shared_world = tables.n_tables == 1
    theta = torch.stack([E.to_device(init(S)) for _ in range(B)])          # [B, S], omega of every candidate
    if per_row:
        for b, o in enumerate(optims):
            o.reset(theta[b])                                               # row views alias theta
    else:
        optims.reset(theta)
    steps = np.zeros(B, dtype=np.int64)
    active = np.arange(B)
    while active.size:
        old = theta.clone()
        compact = shared_world and active.size < B
        if compact:
            idx = torch.as_tensor(active, device=theta.device)
            reward = theta.index_select(0, idx)
            p0_a = p0 if p0_shared else p0.index_select(0, idx)
            ef_a = ef.index_select(0, idx)
        else:
            reward, p0_a, ef_a = theta, p0, ef
        _, grad, _ = E.expected_svf(tables, p0_a, mask, reward, causal=causal, phi=phi,
                                    discount=discount if causal else 0.0, eps_lap=eps_lap, eps_svf=eps_esvf,
                                    e_features=ef_a, fused=None)
        if per_row:
            for j, b in enumerate(active):
                optims[b].step(grad[j] if compact else grad[b])
        else:
            full = torch.zeros_like(theta)
            if compact:
                full.index_copy_(0, idx, grad)
            else:
                full[torch.as_tensor(active, device=theta.device)] = grad[torch.as_tensor(active, device=theta.device)]
            optims.step(full)
        steps[active] += 1
        delta = torch.max(torch.abs(old - theta), dim=1).values.cpu().numpy()   # one host sync per step
        keep = delta[active] > eps
        if max_steps is not None:
            keep &= steps[active] < max_steps
        active = active[keep]
    return theta.clone(), steps
