import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = 32; S = n * n; B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
t = E.gridworld_tables(n, 0.1 + 0.2 * np.arange(B) / B)
r = np.full((B, S), -0.1) + 0.01 * np.random.default_rng(7).standard_normal((B, S)); r[:, S - 1] = 1.0
phi = E.terminal_phi([S - 1], S); rd = E.to_device(r)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    pol = E.soft_vi(t, phi, rd, 0.9)
    torch.cuda.synchronize(); dt = time.time() - t0
c = E.last_info.counts()
print("B=%d soft-VI %.2f ms, mean sweeps %.0f -> %.2f us per world-sweep per SM" % (B, dt * 1e3, c.mean(), dt * 1e6 * 148 / c.sum()))
