"""HBM-bound regime: one large grid in cooperative-grid (streamed) mode, fixed sweep budgets."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
S = n * n
t = E.gridworld_tables(n, 0.2)
p0 = torch.zeros(S, dtype=torch.float64, device="cuda"); p0[0] = 1.0
r = torch.full((S,), -0.1, dtype=torch.float64, device="cuda"); r[S - 1] = 1.0
mask = torch.zeros(S, dtype=torch.uint8, device="cuda"); mask[S - 1] = 1
phi = torch.full((S,), -float("inf"), dtype=torch.float64, device="cuda"); phi[S - 1] = 0.0
for nsw in (20, 60):
    torch.cuda.synchronize(); t0 = time.time()
    pol = E.soft_vi(t, phi, r, 0.9, max_sweeps=nsw, mode=E.MODE_GRID); torch.cuda.synchronize(); t1 = time.time()
    print("soft-VI %d sweeps: %.1f us/sweep -> %.0f GB/s algorithmic (216 B/state)" % (nsw, 1e6 * (t1 - t0) / nsw, 216.0 * S * nsw / (t1 - t0) / 1e9))
for nsw in (50, 200):
    torch.cuda.synchronize(); t0 = time.time()
    d = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=nsw, mode=E.MODE_GRID); torch.cuda.synchronize(); t1 = time.time()
    print("SVF %d sweeps: %.1f us/sweep -> %.0f GB/s algorithmic (84 B/state)" % (nsw, 1e6 * (t1 - t0) / nsw, 84.0 * S * nsw / (t1 - t0) / 1e9))
