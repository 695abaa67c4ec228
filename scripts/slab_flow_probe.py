"""Single-GPU timing of the slab kernels on one n x n world (world size 1): dataflow kernel
(neighbour flags, chunked stop rule) vs the barrier-per-sweep kernel, fixed sweep budgets."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import slab, _irlb200 as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lap_b = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fw_b = int(sys.argv[3]) if len(sys.argv) > 3 else 600
S = n * n
r = np.full(S, -0.1); r[S - 1] = 1.0
phi = np.full(S, -np.inf); phi[S - 1] = 0.0
p0 = np.zeros(S); p0[0] = 1.0
uniform = torch.full((S, 4), 0.25, dtype=torch.float64, device="cuda")
for flow in (0, 1):
    for per_sm in ((0,) if not flow else (0, 2, 1)):
        for chunk in ((0,) if not flow else (32, 64)):
            if per_sm: os.environ["IRLB200_FLOW_CTAS_PER_SM"] = str(per_sm)
            else: os.environ.pop("IRLB200_FLOW_CTAS_PER_SM", None)
            g = slab.PeerSlabGrid(n, 0.2, flow=bool(flow), chunk_sweeps=chunk)
            for rep in range(2):
                E.launch_log = []
                pol, v = g.soft_vi(r, phi, 0.9, 1e-5, max_sweeps=lap_b)
                nl = g.last_n_iter
                d = g.svf(p0, [S - 1], uniform, 1e-5, max_sweeps=fw_b)
                nf = g.last_n_iter
                torch.cuda.synchronize()
                log, E.launch_log = E.launch_log, None
            ms = [a.elapsed_time(b) for nm, a, b in log if nm == "slab_persistent"]
            print("n=%d flow=%d ctas/sm=%s chunk=%s: soft-VI %d sweeps %.2f us/sweep; forward %d sweeps %.2f us/sweep; sum=%.6f"
                  % (n, flow, per_sm or "max", chunk or "-", nl, 1e3 * ms[0] / nl, nf, 1e3 * ms[1] / nf, float(d.sum())), flush=True)
            g.close()
