"""Multi-GPU check of slab mode under torchrun (NCCL): N ranks vs the sparse oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/slab_multi_gpu_check.py [size]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch, torch.distributed as dist
import slab
from oracle import sparse_port as SP

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = n * n
mode = sys.argv[2] if len(sys.argv) > 2 else "nccl"
g = slab.PeerSlabGrid(n, 0.2, icy=True) if mode == "peer" else slab.SlabGrid(n, 0.2, icy=True, chunk=32)
r = np.full(S, -0.1); r[S - 1] = 1.0
phi = np.full(S, -np.inf); phi[S - 1] = 0.0
p0 = np.zeros(S); p0[0] = 1.0
torch.cuda.synchronize(); dist.barrier(); t = time.time()
pol, v = g.soft_vi(g.local(r), g.local(phi), 0.9, 1e-5)
torch.cuda.synchronize(); t_lap = time.time() - t
n_lap = g.last_n_iter
budget = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
t = time.time()
d = g.svf(g.local(p0), [S - 1], pol, 1e-5, max_sweeps=budget)
torch.cuda.synchronize(); t_svf = time.time() - t
n_svf = g.last_n_iter
pol_full, d_full = g.gather(pol).cpu().numpy(), g.gather(d).cpu().numpy()
if rank == 0:
    print("mode=%s " % mode, end="")
    print("ranks=%d n=%d soft-VI %d sweeps %.3f s (%.1f us/sweep); SVF %d sweeps %.3f s (%.1f us/sweep)"
          % (world, n, n_lap, t_lap, 1e6 * t_lap / n_lap, n_svf, t_svf, 1e6 * t_svf / n_svf))
    if n <= 128:
        mdp = SP.icy_gridworld_sparse(n, 0.2)
        pa, k = SP.local_causal_action_probabilities(mdp, [S - 1], r, 0.9)
        dref, _ = SP.expected_svf_from_policy(mdp, p0, [S - 1], pa, 1e-5, max_sweeps=budget)
        assert k == n_lap, (k, n_lap)
        np.testing.assert_allclose(pol_full, pa, rtol=1e-10)
        np.testing.assert_allclose(d_full, dref, rtol=1e-10, atol=1e-300)
        print("slab multi-GPU parity OK (policy, SVF to 1e-10; %d soft-VI sweeps identical)" % k)
if mode == "peer":
    g.close()
dist.destroy_process_group()
