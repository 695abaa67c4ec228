"""One forward launch on identical worlds (for ncu): args n B sweeps"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n, B, sweeps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
S = n * n
p0 = np.zeros(S); p0[0] = 1.0
mask = E.terminal_mask([S - 1], S)
tabs = E.gridworld_tables(n, np.full(B, 0.2))
r = np.tile(-np.log(4.0) + 0.01 * np.random.default_rng(1).standard_normal(S), (B, 1))
pol = E.backward(tabs, mask, E.to_device(r))
d = E.svf(tabs, p0, mask, pol, 1e-5, max_sweeps=sweeps)
torch.cuda.synchronize()
print(E.last_info.counts()[:3])
