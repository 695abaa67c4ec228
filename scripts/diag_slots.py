"""4-slot vs 5-slot tables in cooperative-grid mode: which outputs differ, and by how much."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
S = n * n
r = np.full(S, -0.1); r[S - 1] = 1.0
p0 = np.zeros(S); p0[0] = 1.0
mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
for streamed in ("0", "1"):
    os.environ["IRLB200_FORCE_STREAMED"] = streamed
    out = []
    for slots in (5, 4):
        t = E.gridworld_tables(n, 0.2, slots=slots)
        pol = E.soft_vi(t, phi, r, 0.9, mode=E.MODE_GRID, want_value=True)
        pol, val = pol if isinstance(pol, tuple) else (pol, None)
        n_lap = int(E.last_info.counts()[0])
        d = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=3000, mode=E.MODE_GRID)
        n_svf = int(E.last_info.counts()[0])
        v = E.value_iteration(t, r, 0.9, 1e-4, mode=E.MODE_GRID)
        n_vi = int(E.last_info.counts()[0])
        pb = E.backward(t, mask, np.full(S, -np.log(4.0)), n_sweeps=200, mode=E.MODE_GRID)
        out.append(dict(pol=pol, val=val, d=d, v=v, pb=pb, counts=(n_lap, n_svf, n_vi)))
    print("streamed", streamed, "counts", out[0]["counts"], out[1]["counts"])
    for k in ("pol", "val", "d", "v", "pb"):
        a, b = out[0][k], out[1][k]
        if a is None: continue
        a, b = a.cpu().numpy().ravel(), b.cpu().numpy().ravel()
        ne = np.flatnonzero(a != b)
        print("  %-4s differing entries %d / %d   max rel %.3g   first %s" % (
            k, ne.size, a.size, np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)) if ne.size else 0.0, ne[:5]))
