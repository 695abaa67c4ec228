N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
IRLB200_SLAB_FLOW=1 timeout 120 $TR scripts/slab_multi_gpu_check.py 128 peer 3000 2>&1 | grep "mode=\|parity\|rror"
IRLB200_SLAB_FLOW=0 timeout 200 $TR scripts/slab_c5_bench.py 2048 300 2000 2>&1 | grep "^C5.*peer"
IRLB200_SLAB_FLOW=1 timeout 200 $TR scripts/slab_c5_bench.py 2048 300 2000 2>&1 | grep "^C5.*peer\|rror"
IRLB200_SLAB_FLOW=1 IRLB200_FLOW_CTAS_PER_SM=2 timeout 200 $TR scripts/slab_c5_bench.py 2048 300 2000 2>&1 | grep "^C5.*peer\|rror"
