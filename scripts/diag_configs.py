"""Timing of the non-headline configs: C1/C2 (5x5 irl / irl_causal end to end), C3 (128x128 single MDP)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _irlb200 as E, maxent as M, gridworld as W, optimizer as O, solver as S
from test_oracle_golden import load_trajectories
g = np.load(os.path.join(ROOT, "tests/golden/e2e_5x5.npz"))
world = W.IcyGridWorld(5, 0.2); tjs = load_trajectories(g); F = W.state_features(world)
class C:
    def __init__(s, i): s.i, s.n = i, 0
    def reset(s, p): s.i.reset(p)
    def step(s, gr): s.n += 1; return s.i.step(gr)
for name, fn in (("C1 irl", lambda o: M.irl(world.p_transition, F, [24], tjs, o, O.Constant(1.0))),
                 ("C2 irl_causal g=0.9", lambda o: M.irl_causal(world.p_transition, F, [24], tjs, o, O.Constant(1.0), 0.9))):
    for rep in range(2):
        o = C(O.ExpSga(lr=O.linear_decay(lr0=0.2)))
        torch.cuda.synchronize(); t = time.time(); r = fn(o); torch.cuda.synchronize(); dt = time.time() - t
    print("%s: %d steps in %.3f s -> %.0f grad-steps/s" % (name, o.n, dt, o.n / dt))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
Sn = n * n
t = E.gridworld_tables(n, 0.2)
p0 = np.zeros(Sn); p0[0] = 1.0
mask, phi = E.terminal_mask([Sn - 1], Sn), E.terminal_phi([Sn - 1], Sn)
rg = np.full(Sn, -0.1); rg[Sn - 1] = 1.0
for mode, nm in ((E.MODE_GRID, "grid"), (E.MODE_CLUSTER, "cluster")):
    torch.cuda.synchronize(); t0 = time.time()
    pol = E.soft_vi(t, phi, rg, 0.9, mode=E.MODE_GRID); torch.cuda.synchronize(); t1 = time.time()
    nl = E.last_info.counts()[0]
    d = E.svf(t, p0, mask, pol[0], 1e-5, max_sweeps=400000, mode=mode); torch.cuda.synchronize(); t2 = time.time()
    ns = E.last_info.counts()[0]
    print("C3 %dx%d %s: soft-VI %d sweeps %.3f s (%.2f us/sweep); SVF %d sweeps %.3f s (%.2f us/sweep) status %d"
          % (n, n, nm, nl, t1 - t0, 1e6 * (t1 - t0) / nl, ns, t2 - t1, 1e6 * (t2 - t1) / ns, E.last_info.stati()[0]))
rr = -np.log(4.0) + 0.01 * np.random.default_rng(0).standard_normal(Sn)
torch.cuda.synchronize(); t0 = time.time()
pb = E.backward(t, mask, rr, mode=E.MODE_GRID); torch.cuda.synchronize(); t1 = time.time()
print("C3 backward %d sweeps %.3f s (%.2f us/sweep)" % (2 * Sn, t1 - t0, 1e6 * (t1 - t0) / (2 * Sn)))
