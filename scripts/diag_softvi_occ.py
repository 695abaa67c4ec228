import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/irl-maxent_b200")
import torch
import _irlb200 as E
n = 2048; S = n * n
dev = "cuda"
r = torch.full((S,), -0.1, dtype=torch.float64, device=dev); r[S - 1] = 1.0
phi = torch.full((S,), -float("inf"), dtype=torch.float64, device=dev); phi[S - 1] = 0.0
t = E.gridworld_tables(n, 0.2, slots=4)
ref = None
for rep in range(3):
    for occ in ("0", "3"):
        os.environ["IRLB200_SOFTVI_OCC"] = occ
        E.launch_log = []
        pol = E.soft_vi(t, phi, r, 0.9, max_sweeps=150, mode=E.MODE_GRID)
        torch.cuda.synchronize()
        log, E.launch_log = E.launch_log, None
        ms = {name: a.elapsed_time(b) for name, a, b in log}
        if ref is None: ref = pol.clone()
        if rep: print("occ", occ, "soft-VI %.1f us/sweep" % (1e3 * ms["soft_vi"] / 150), "same", bool((pol == ref).all()), flush=True)
