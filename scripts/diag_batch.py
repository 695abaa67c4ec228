"""Diagnostic: sweep-count distribution and kernel time of a batch of C4 worlds."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E, maxent as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
maxs = int(sys.argv[4]) if len(sys.argv) > 4 else 1000000
S = n * n
ps = 0.1 + 0.2 * (np.arange(B) * (4096 // B)) / 4096.0
tabs = E.gridworld_tables(n, ps)
rng = np.random.default_rng(1000)
r = -np.log(4.0) + noise * rng.standard_normal((B, S))
p0 = np.zeros(S); p0[0] = 1.0
mask = E.terminal_mask([S - 1], S)
rd = E.to_device(r)
for it in range(2):
    torch.cuda.synchronize(); t = time.time()
    pol = E.backward(tabs, mask, rd)
    torch.cuda.synchronize(); tb = time.time() - t
    t = time.time()
    d = E.svf(tabs, p0, mask, pol, 1e-5, max_sweeps=maxs)
    torch.cuda.synchronize(); tf = time.time() - t
    c = E.last_info.counts(); st = E.last_info.stati()
    print("B=%d n=%d backward %.1f ms, svf %.1f ms; sweeps min/med/mean/max %d/%d/%d/%d; capped %d; sum sweeps %.3g; ns/sweep/world-on-148 %.1f"
          % (B, n, tb * 1e3, tf * 1e3, c.min(), np.median(c), c.mean(), c.max(), (st == 2).sum(), c.sum(),
             tf * 1e9 / (c.sum() / 148.0)))
print(np.sort(c)[-10:])
