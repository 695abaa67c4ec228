TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for n in 64 128; do
  IRLB200_SLAB_FLOW=1 timeout 120 $TR scripts/slab_multi_gpu_check.py $n peer 3000 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|W1018" 
done
IRLB200_SLAB_FLOW=0 timeout 200 $TR scripts/slab_c5_bench.py 2048 300 600 2>&1 | grep "^C5"
IRLB200_SLAB_FLOW=1 timeout 200 $TR scripts/slab_c5_bench.py 2048 300 600 2>&1 | grep "^C5\|rror"
