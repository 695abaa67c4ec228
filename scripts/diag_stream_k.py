"""2048x2048 streamed cooperative-grid kernels: 4-slot vs 5-slot tables, alternating, warmed up."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import torch
import _irlb200 as E
n = 2048; S = n * n
dev = "cuda"
p0 = torch.zeros(S, dtype=torch.float64, device=dev); p0[0] = 1.0
r = torch.full((S,), -0.1, dtype=torch.float64, device=dev); r[S - 1] = 1.0
mask = torch.zeros(S, dtype=torch.uint8, device=dev); mask[S - 1] = 1
phi = torch.full((S,), -float("inf"), dtype=torch.float64, device=dev); phi[S - 1] = 0.0
tabs = {k: E.gridworld_tables(n, 0.2, slots=k) for k in (5, 4)}
for rep in range(4):
    for k in (5, 4, 4, 5):
        E.launch_log = []
        pol = E.soft_vi(tabs[k], phi, r, 0.9, max_sweeps=150, mode=E.MODE_GRID)
        v = E.value_iteration(tabs[k], r, 0.9, 1e-30, max_sweeps=300, mode=E.MODE_GRID)
        d = E.svf(tabs[k], p0, mask, pol, 1e-5, max_sweeps=400, mode=E.MODE_GRID)
        torch.cuda.synchronize()
        log, E.launch_log = E.launch_log, None
        ms = {name: a.elapsed_time(b) for name, a, b in log}
        if rep:
            print("rep %d slots %d: soft-VI %.1f us/sweep, VI %.1f us/sweep, forward %.1f us/sweep" % (
                rep, k, 1e3 * ms["soft_vi"] / 150, 1e3 * ms["value_iteration"] / 300, 1e3 * ms["svf"] / 400), flush=True)
