"""128x128 MaxEnt step: backward pass in cluster (push) mode vs cooperative grid, forward pass."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = 128; S = n * n
t = E.gridworld_tables(n, 0.2)
r = E.to_device(-np.log(4.0) + 0.01 * np.random.default_rng(0).standard_normal(S))
mask = E.terminal_mask([S - 1], S)
p0 = np.zeros(S); p0[0] = 1.0
for name, mode, cs in (("grid", E.MODE_GRID, 0), ("cluster8", E.MODE_CLUSTER, 8), ("cluster16", E.MODE_CLUSTER, 16)):
    if cs: os.environ["IRLB200_CLUSTER_SIZE"] = str(cs)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pol = E.backward(t, mask, r, mode=mode)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("backward %-9s %.2f ms = %.3f us per sweep" % (name, dt * 1e3, dt * 1e6 / (2 * S)), flush=True)
os.environ.pop("IRLB200_CLUSTER_SIZE")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = E.svf(t, p0, mask, pol, 1e-5)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
c = int(E.last_info.counts().ravel()[0])
print("forward: %.2f ms, %d sweeps, %.3f us per sweep" % (dt * 1e3, c, dt * 1e6 / c))
