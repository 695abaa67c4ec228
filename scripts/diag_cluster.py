"""128x128 forward pass: cluster (barrier.cluster) vs push (st.async + mbarrier) variants, cluster sizes 8 / 16.
Checks bitwise equality against the cooperative-grid kernel and prints us per sweep."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import _irlb200 as E

def run(n, budget=None):
    S = n * n
    t = E.gridworld_tables(n, 0.2)
    r = np.full(S, -0.1); r[S - 1] = 1.0
    p0 = np.zeros(S); p0[0] = 1.0
    mask, phi = E.terminal_mask([S - 1], S), E.terminal_phi([S - 1], S)
    pol = E.soft_vi(t, phi, r, 0.9)
    os.environ.pop("IRLB200_CLUSTER_SIZE", None)
    d_g = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=budget, mode=E.MODE_GRID)
    n_g = int(E.last_info.counts()[0][0] if E.last_info.counts().ndim > 1 else E.last_info.counts()[0])
    for push in (0, 1):
        for cs in (2, 4, 8, 16):
            os.environ["IRLB200_CLUSTER_PUSH"] = str(push)
            os.environ["IRLB200_CLUSTER_SIZE"] = str(cs)
            try:
                best = None
                for rep in range(3):
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    d = E.svf(t, p0, mask, pol, 1e-5, max_sweeps=budget, mode=E.MODE_CLUSTER)
                    torch.cuda.synchronize(); dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                cnt = E.last_info.counts().ravel()[0]; stt = E.last_info.stati().ravel()[0]
                print("n=%d push=%d cluster=%2d: %8.2f ms  sweeps %d status %d  %.3f us/sweep  bitwise=%s count_ok=%s" % (
                    n, push, cs, best * 1e3, cnt, stt, 1e6 * best / max(int(cnt), 1), bool((d == d_g).all()), int(cnt) == n_g), flush=True)
            except Exception as e:
                print("n=%d push=%d cluster=%2d: %s" % (n, push, cs, str(e)[:100]), flush=True)

for n, b in ((16, None), (64, None), (128, 30000), (128, None)):
    run(n, b)
