timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.log 2> gpurun_out/r2_bench_n8.err
tail -c 800 gpurun_out/r2_bench_n8.err
python scripts/_show_oc.py gpurun_out/r2_bench_n8.log 6000
