import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import torch, torch.distributed as dist
import bench
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    out = bench.c5_slab(dist.get_world_size(), dist.get_rank(), dev)
    if dist.get_rank() == 0:
        print(json.dumps(out, indent=1))
except Exception:
    import traceback
    sys.stderr.write("[rank %d]\n%s\n" % (dist.get_rank(), traceback.format_exc()))
dist.destroy_process_group()
