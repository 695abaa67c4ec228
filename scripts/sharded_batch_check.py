"""Multi-GPU check of sharding.compute_expected_svf_batch_sharded under torchrun (NCCL): the gathered batch
equals the single-rank batch bitwise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import numpy as np, torch, torch.distributed as dist
import sharding, _irlb200 as E, maxent as M
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
B, n = 37, 16
S = n * n
ps = 0.1 + 0.2 * np.arange(B) / B
r = -np.log(4.0) + 0.01 * np.random.default_rng(3).standard_normal((B, S))
p0 = np.zeros(S); p0[0] = 1.0
ef = np.random.default_rng(4).random((B, S))
d, g, (b0, b1) = sharding.compute_expected_svf_batch_sharded(n, ps, p0, [S - 1], r, e_features=ef, fused=False)
assert d.shape == (B, S) and g.shape == (B, S)
if rank == 0:
    d1, g1 = M.compute_expected_svf_batch(E.gridworld_tables(n, ps), p0, [S - 1], r, e_features=ef, fused=False)
    assert (d == d1).all() and (g == g1).all()
    print("sharded batch over %d ranks == single-rank batch (bitwise); rank 0 owned [%d, %d)" % (world, b0, b1))
dist.destroy_process_group()
