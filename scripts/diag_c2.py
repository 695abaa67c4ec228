import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _irlb200 as E, maxent as M, gridworld as W, optimizer as O
from test_oracle_golden import load_trajectories
g = np.load(os.path.join(ROOT, "tests/golden/e2e_5x5.npz"))
world = W.IcyGridWorld(5, 0.2); tjs = load_trajectories(g); F = W.state_features(world)
for causal in (False, True):
    for rep in range(2):
        E.launch_log = []
        o = O.ExpSga(lr=O.linear_decay(lr0=0.2))
        torch.cuda.synchronize(); t = time.time()
        r = M.irl_causal(world.p_transition, F, [24], tjs, o, O.Constant(1.0), 0.9) if causal else M.irl(world.p_transition, F, [24], tjs, o, O.Constant(1.0))
        torch.cuda.synchronize(); dt = time.time() - t
        log, E.launch_log = E.launch_log, None
    k = sum(a.elapsed_time(b) for _, a, b in log)
    print("causal=%s steps=%d wall %.1f ms, kernel (events) %.1f ms = %.3f ms/step, host overhead %.3f ms/step"
          % (causal, len(log), dt * 1e3, k, k / len(log), (dt * 1e3 - k) / len(log)))
