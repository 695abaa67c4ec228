"""Timing-only variants of the push kernel (IRLB200_PUSH_VAR bits; 2, 8, 10 give wrong results by design)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
for n, cs in ((16, 2), (64, 8), (128, 8), (128, 16)):
    S = n * n
    t = E.gridworld_tables(n, 0.2)
    r = np.full(S, -0.1); r[S - 1] = 1.0
    p0 = np.zeros(S); p0[0] = 1.0
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    pol = E.soft_vi(t, phi, r, 0.9)
    os.environ["IRLB200_CLUSTER_SIZE"] = str(cs)
    for var in (0, 1, 2, 4, 8, 10):
        os.environ["IRLB200_PUSH_VAR"] = str(var)
        best = None
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            d = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=3000, mode=E.MODE_CLUSTER)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        cnt = int(E.last_info.counts().ravel()[0])
        print("n=%3d cluster=%2d var=%2d: %.3f us/sweep (%d sweeps)" % (n, cs, var, 1e6 * best / cnt, cnt), flush=True)
