TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu -k "slab" 2>&1 | tail -3
timeout 200 $TR scripts/slab_flow_probe.py 1024 600 20000 IRLB200_FLOW_RESIDENT=0 IRLB200_FLOW_RESIDENT=1 2>&1 | grep "^n=\|rror\|abort" | head -5
timeout 200 $TR scripts/slab_flow_probe.py 2048 0 3000 IRLB200_FLOW_RESIDENT=1 2>&1 | grep "^n=\|rror\|abort" | head -5
