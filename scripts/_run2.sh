TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for rep in 1 2 3; do
timeout 100 $TR scripts/slab_flow_probe.py 2048 0 6000 IRLB200_SLAB_FLOW=1 IRLB200_FLOW_CHUNK=64 2>&1 | grep "^n=\|aborted" | sort | uniq | head -5
done
timeout 100 $TR scripts/slab_flow_probe.py 1024 600 30000 IRLB200_SLAB_FLOW=1 2>&1 | grep "^n=\|aborted" | sort | uniq | head -5
timeout 400 $TR scripts/_c5_only.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -70
