"""C5: one 2048x2048 IcyGridWorld, row slabs over N GPUs, persistent peer-store kernels vs the
NCCL baseline; fixed sweep budgets (kernel time from CUDA events on each rank, max over ranks)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch, torch.distributed as dist
import slab, _irlb200 as E

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lap_budget = int(sys.argv[2]) if len(sys.argv) > 2 else 300
fw_budget = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
S = n * n
pol_keep = None
for mode in ("peer", "nccl"):
    g = slab.PeerSlabGrid(n, 0.2) if mode == "peer" else slab.SlabGrid(n, 0.2, chunk=50)
    r = np.full(g.cnt, -0.1)
    if g.hi == S: r[-1] = 1.0
    phi = np.full(g.cnt, -np.inf)
    if g.hi == S: phi[-1] = 0.0
    p0 = np.zeros(g.cnt)
    if g.lo == 0: p0[0] = 1.0
    lb = None if mode == "peer" else 100                 # peer: to convergence (exact count); baseline: timing only
    fb = fw_budget if mode == "peer" else min(fw_budget, 200)
    for rep in range(2):                                  # rep 0 warms up NCCL P2P, workspaces, clocks
        E.launch_log = []
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        pol, v = g.soft_vi(r, phi, 0.9, 1e-5, max_sweeps=lb)
        n_lap = g.last_n_iter
        torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
        if mode == "peer":
            pol_keep = pol
        d = g.svf(p0, [S - 1], pol_keep, 1e-5, max_sweeps=fb)
        n_fw = g.last_n_iter
        torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
        log, E.launch_log = E.launch_log, None
    if mode == "peer":
        k = [a.elapsed_time(b) / 1e3 for nm, a, b in log if nm == "slab_persistent"]
        tl = torch.tensor(k[:2], device="cuda")
    else:
        tl = torch.tensor([t1 - t0, t2 - t1], device="cuda")
    dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    if rank == 0:
        tl = tl.tolist()
        print("C5 %dx%d ranks=%d mode=%s (%s): soft-VI %d sweeps %.1f us/sweep (%.0f GB/s aggregate algorithmic); "
              "forward %d sweeps %.1f us/sweep (%.0f GB/s); sum(svf) local %.6g" % (
                  n, n, world, mode, "kernel time, CUDA events" if mode == "peer" else "wall time incl. collectives",
                  n_lap, 1e6 * tl[0] / n_lap, 216.0 * S * n_lap / tl[0] / 1e9, n_fw, 1e6 * tl[1] / n_fw,
                  84.0 * S * n_fw / tl[1] / 1e9, float(d.sum())))
    if mode == "peer":
        g.close()
dist.destroy_process_group()
