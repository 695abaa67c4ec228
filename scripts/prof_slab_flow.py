"""Short single-GPU run of the dataflow slab kernel for ncu: 724x724 (the per-GPU slab size of 2048x2048 on
8 GPUs), forward pass, fixed sweep budget."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import slab
n = 724; S = n * n
g = slab.PeerSlabGrid(n, 0.2, flow=True)
p0 = np.zeros(S); p0[0] = 1.0
uniform = torch.full((S, 4), 0.25, dtype=torch.float64, device="cuda")
for _ in range(2):
    d = g.svf(p0, [S - 1], uniform, 1e-5, max_sweeps=int(sys.argv[1]) if len(sys.argv) > 1 else 256)
print(g.last_n_iter, float(d.sum()))
g.close()
