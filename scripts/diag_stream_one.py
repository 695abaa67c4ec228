"""One streamed cooperative-grid forward launch at 2048x2048 (for ncu): fixed 100 sweeps.
argv[1]: table slots (5 or 4, default 4)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import torch
import _irlb200 as E
n = 2048; S = n * n
t = E.gridworld_tables(n, 0.2, slots=int(sys.argv[1]) if len(sys.argv) > 1 else 4)
p0 = torch.zeros(S, dtype=torch.float64, device="cuda"); p0[0] = 1.0
mask = torch.zeros(S, dtype=torch.uint8, device="cuda"); mask[S - 1] = 1
pol = torch.full((1, S, 4), 0.25, dtype=torch.float64, device="cuda")
d = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=100, mode=E.MODE_GRID)
torch.cuda.synchronize()
print(E.last_info.counts())
