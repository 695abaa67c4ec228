"""Experiment: 4-slot tables (drop the always-empty 5th slot) in the streamed cooperative-grid kernels."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = 2048; S = n * n
t5 = E.gridworld_tables(n, 0.2)
assert float(t5.succ_p[:, :, 4, :].abs().max()) == 0.0 and float(t5.pred_p[:, :, 4, :].abs().max()) == 0.0
t4 = E.Tables(S, 4, 4, 4, t5.succ_idx[:, :4, :].contiguous(), t5.succ_p[:, :, :4, :].contiguous(),
              t5.pred_idx[:, :4, :].contiguous(), t5.pred_p[:, :, :4, :].contiguous(), 1, 0)
p0 = torch.zeros(S, dtype=torch.float64, device="cuda"); p0[0] = 1.0
r = torch.full((S,), -0.1, dtype=torch.float64, device="cuda"); r[S - 1] = 1.0
mask = torch.zeros(S, dtype=torch.uint8, device="cuda"); mask[S - 1] = 1
phi = torch.full((S,), -float("inf"), dtype=torch.float64, device="cuda"); phi[S - 1] = 0.0
pol = torch.full((1, S, 4), 0.25, dtype=torch.float64, device="cuda")
res = {}
for name, t in (("K=5 static", t5), ("K=4 dynamic", t4)):
    for rep in range(2):
        E.launch_log = []
        v = E.soft_vi(t, phi, r, 0.9, max_sweeps=150, mode=E.MODE_GRID, want_value=True)[1]
        d = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=400, mode=E.MODE_GRID)
        torch.cuda.synchronize()
        log, E.launch_log = E.launch_log, None
    ms = {nm: a.elapsed_time(b) for nm, a, b in log}
    res[name] = (v, d)
    print("%s: soft-VI %.1f us/sweep, forward %.1f us/sweep" % (name, 1e3 * ms["soft_vi"] / 150, 1e3 * ms["svf"] / 400))
print("values equal:", bool((res["K=5 static"][0] == res["K=4 dynamic"][0]).all()), bool((res["K=5 static"][1] == res["K=4 dynamic"][1]).all()))
