"""Timing of the slab kernels on one n x n world (any world size; single GPU works): fixed sweep budgets,
kernel time from CUDA events (max over ranks).  Environment variants are given as KEY=VAL,KEY=VAL groups:
    python scripts/slab_flow_probe.py 724 150 2000 IRLB200_SLAB_FLOW=0 IRLB200_FLOW_FWD=14 IRLB200_FLOW_FWD=42
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch
import slab, _irlb200 as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lap_b = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fw_b = int(sys.argv[3]) if len(sys.argv) > 3 else 600
variants = sys.argv[4:] or ["IRLB200_SLAB_FLOW=0", "IRLB200_SLAB_FLOW=1"]
dist = None
if "RANK" in os.environ:
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if dist else 0
world = dist.get_world_size() if dist else 1
S = n * n
for var in variants:
    kv = dict(x.split("=") for x in var.split(",") if x)
    saved = {k: os.environ.get(k) for k in kv}
    os.environ.update(kv)
    g = slab.PeerSlabGrid(n, 0.2)
    r = np.full(g.cnt, -0.1); phi = np.full(g.cnt, -np.inf); p0 = np.zeros(g.cnt)
    if g.hi == S: r[-1] = 1.0; phi[-1] = 0.0
    if g.lo == 0: p0[0] = 1.0
    uniform = torch.full((g.cnt, 4), 0.25, dtype=torch.float64, device="cuda")
    for rep in range(2):
        E.launch_log = []
        nl = 0
        if lap_b > 0:
            pol, v = g.soft_vi(r, phi, 0.9, 1e-5, max_sweeps=lap_b)
            nl = g.last_n_iter
        d = g.svf(p0, [S - 1], uniform, 1e-5, max_sweeps=fw_b)
        nf = g.last_n_iter
        torch.cuda.synchronize()
        log, E.launch_log = E.launch_log, None
    ms = [a.elapsed_time(b) for nm, a, b in log if nm == "slab_persistent"]
    t = torch.tensor(ms, device="cuda", dtype=torch.float64)
    tot = d.sum().reshape(1).clone()
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms = t.tolist()
    if rank == 0:
        lap = ("soft-VI %d sweeps %.2f us/sweep; " % (nl, 1e3 * ms[0] / nl)) if lap_b > 0 else ""
        print("n=%d ranks=%d [%s]: %sforward %d sweeps %.2f us/sweep; sum(svf)=%.9f"
              % (n, world, var, lap, nf, 1e3 * ms[-1] / nf, float(tot)), flush=True)
    g.close()
    for k, v in saved.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
if dist:
    dist.destroy_process_group()
