timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/_c5_only.py > gpurun_out/r2_c5_n8.log 2>&1
python - <<'PY'
import json,re
t=open("gpurun_out/r2_c5_n8.log").read()
i=t.index("{\n")
d=json.loads(t[i:t.rindex("}")+1])
print(json.dumps(d["dataflow_kernel"])[:700]); print(d["parity"])
PY
