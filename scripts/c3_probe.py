"""C3 (one 128x128 world): us per sweep of the forward / backward cluster kernels and the soft-VI kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = 128; S = n * n
tabs = E.gridworld_tables(n, 0.2)
p0 = np.zeros(S); p0[0] = 1.0
r = np.full(S, -0.1); r[S - 1] = 1.0
mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
rd, p0d = E.to_device(r), E.to_device(p0)
for rep in range(3):
    E.launch_log = []
    pol = E.soft_vi(tabs, phi, rd, 0.9, mode=E.MODE_AUTO)
    nl = int(E.last_info.n_iter.item())
    d = E.svf(tabs, p0d, mask, pol, 1e-5)
    nf = int(E.last_info.n_iter.item())
    rm = E.to_device(-np.log(4.0) + 0.01 * np.random.default_rng(0).standard_normal(S))
    pb = E.backward(tabs, mask, rm)
    torch.cuda.synchronize()
    log, E.launch_log = E.launch_log, None
ms = {nm: a.elapsed_time(b) for nm, a, b in log}
print("C3 128x128: soft-VI %d sweeps %.3f us/sweep; forward %d sweeps %.4f us/sweep (%.1f ms); backward %d sweeps %.4f us/sweep"
      % (nl, 1e3 * ms["soft_vi"] / nl, nf, 1e3 * ms["svf"] / nf, ms["svf"], 2 * S, 1e3 * ms["backward"] / (2 * S)))
