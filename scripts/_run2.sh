TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for n in 64 128; do
  timeout 120 $TR scripts/slab_multi_gpu_check.py $n peer 3000 2>&1 | grep "mode=\|parity\|rror" 
done
timeout 300 $TR scripts/slab_flow_probe.py 1024 1200 3000 IRLB200_SLAB_FLOW=0 IRLB200_FLOW_FWD=14 IRLB200_FLOW_FWD=24 IRLB200_FLOW_FWD=42 IRLB200_FLOW_EDGE_CTAS=0 IRLB200_FLOW_EDGE_CTAS=8 2>&1 | grep "^n=\|rror"
timeout 200 $TR scripts/slab_c5_bench.py 2048 300 600 2>&1 | grep "^C5.*peer\|rror"
