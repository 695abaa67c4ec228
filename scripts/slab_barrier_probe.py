"""Per-sweep cost of the persistent slab kernel on a tiny grid (sync-dominated): kernel time from CUDA events."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch, torch.distributed as dist
import slab, _irlb200 as E
multi = "RANK" in os.environ
if multi:
    local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if multi else 0
world = dist.get_world_size() if multi else 1
for n in [int(a) for a in sys.argv[1:]] or [64]:
    S = n * n
    g = slab.PeerSlabGrid(n, 0.2)
    p0 = np.zeros(g.cnt)
    if g.lo == 0: p0[0] = 1.0
    pol = np.full((g.cnt, 4), 0.25)
    for budget in (2000, 6000):
        E.launch_log = []
        d = g.svf(p0, [S - 1], pol, 1e-5, max_sweeps=budget)
        log, E.launch_log = E.launch_log, None
        ms = [a.elapsed_time(b) for nm, a, b in log if nm == "slab_persistent"][0]
        if rank == 0:
            print("n=%d ranks=%d forward %d sweeps: kernel %.2f ms -> %.2f us/sweep" % (n, world, g.last_n_iter, ms, 1e3 * ms / g.last_n_iter))
    g.close()
if multi:
    dist.destroy_process_group()
