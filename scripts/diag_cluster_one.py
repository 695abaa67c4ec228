"""One cluster-mode forward pass (for ncu): n, cluster size and sweep budget from argv."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
os.environ["IRLB200_CLUSTER_SIZE"] = sys.argv[2] if len(sys.argv) > 2 else "8"
budget = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
S = n * n
t = E.gridworld_tables(n, 0.2)
r = np.full(S, -0.1); r[S - 1] = 1.0
p0 = np.zeros(S); p0[0] = 1.0
mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
pol = E.soft_vi(t, phi, r, 0.9)
d = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=budget, mode=E.MODE_CLUSTER)
torch.cuda.synchronize()
print("sweeps", int(E.last_info.counts().ravel()[0]), "sum", float(d.sum()))
