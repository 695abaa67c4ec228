"""Tiny invocations of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E, maxent as M
rng = np.random.default_rng(0)
for n, modes in ((5, (E.MODE_AUTO, E.MODE_GRID)), (8, (E.MODE_AUTO, E.MODE_GRID)), (16, (E.MODE_AUTO, E.MODE_CLUSTER))):
    S = n * n
    t = E.gridworld_tables(n, [0.2, 0.3])
    r = -np.log(4.0) + 0.01 * rng.standard_normal((2, S))
    p0 = np.zeros(S); p0[0] = 1.0
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    pol = E.backward(t, mask, r, n_sweeps=12)
    polc = E.soft_vi(t, phi, r, 0.9, max_sweeps=12)
    E.value_iteration(t, r, 0.9, max_sweeps=12)
    for mode in modes:
        if mode == E.MODE_GRID:
            E.svf(t.select(0), p0, mask, pol[0], max_sweeps=12, mode=mode)
        else:
            E.svf(t, p0, mask, pol, max_sweeps=12, mode=mode)
    M.compute_expected_svf_batch(t, p0, [S - 1], r, fused=True, max_sweeps=12)
    M.compute_expected_svf_batch(t, p0, [S - 1], r, causal=True, discount=0.9, fused=True, max_sweeps=12)
P = np.zeros((7, 7, 3))
for s in range(7):
    for a in range(3):
        P[s, (s + a + 1) % 7, a] = 0.6; P[s, 6, a] += 0.4
t = E.compress_dense(P)
E.svf(t, np.full(7, 1 / 7), E.terminal_mask([6], 7), np.full((7, 3), 0.3), max_sweeps=12)
E.soft_vi(t, E.terminal_phi([6], 7), np.zeros(7), 0.8, max_sweeps=12)
torch.cuda.synchronize()
print("sanitize run ok")
