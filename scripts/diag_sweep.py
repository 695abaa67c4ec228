"""Per-sweep latency of the forward kernel: identical worlds, fixed sweep budget."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
S = n * n
p0 = np.zeros(S); p0[0] = 1.0
mask = E.terminal_mask([S - 1], S)
for B in (148, 296, 444, 592, 1184):
    tabs = E.gridworld_tables(n, np.full(B, 0.2))
    r = np.tile(-np.log(4.0) + 0.01 * np.random.default_rng(1).standard_normal(S), (B, 1))
    pol = E.backward(tabs, mask, E.to_device(r))
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        d = E.svf(tabs, p0, mask, pol, 1e-5, max_sweeps=sweeps)
        torch.cuda.synchronize(); tf = time.time() - t
    c = E.last_info.counts()
    per = tf / c[0]
    print("B=%4d (%.1f/SM) sweeps=%d  %.1f ms  -> %.0f ns per sweep-round, %.1f ns per world-sweep per SM"
          % (B, B / 148.0, c[0], tf * 1e3, per * 1e9, tf * 1e9 / (c.sum() / 148.0)))
