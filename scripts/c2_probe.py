"""C1 / C2 end-to-end rates (the bench's side lines) in isolation."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import gridworld as W, maxent as M, optimizer as O, trajectory as T
g = np.load(os.path.join(ROOT, "tests", "golden", "e2e_5x5.npz"))
tjs, off = [], 0
for length in g["traj_len"]:
    tjs.append(T.Trajectory([tuple(int(v) for v in row) for row in g["traj_flat"][off:off + length]])); off += length
world = W.IcyGridWorld(5, 0.2); F = W.state_features(world)
for name, fn in (("C1", lambda o: M.irl(world.p_transition, F, [24], tjs, o, O.Constant(1.0))),
                 ("C2", lambda o: M.irl_causal(world.p_transition, F, [24], tjs, o, O.Constant(1.0), 0.9))):
    best = None
    for _ in range(4):
        o = O.ExpSga(lr=O.linear_decay(lr0=0.2))
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(o); torch.cuda.synchronize()
        dt = time.perf_counter() - t; best = dt if best is None else min(best, dt)
    print("%s: %d steps, %.4f s, %.0f grad-steps/s, reward[24]=%.12f" % (name, o.k, best, o.k / best, r[24]))
