"""One 32x32 world through the public single-world API: fused vs split launches."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = 32; S = n * n
t = E.gridworld_tables(n, 0.2)
r = -np.log(4.0) + 0.01 * np.random.default_rng(1).standard_normal(S)
rg = np.full(S, -0.1); rg[S - 1] = 1.0
p0 = np.zeros(S); p0[0] = 1.0
mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
for fused in (True, False, None):
    for causal in (False, True):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            d, _, _ = E.expected_svf(t, p0, mask, rg if causal else r, causal=causal, phi=phi, discount=0.9, fused=fused)
            torch.cuda.synchronize(); dt = time.time() - t0
        print("fused=%s causal=%s: %.2f ms  sweeps %s" % (fused, causal, dt * 1e3, E.last_info.counts()[0]))
