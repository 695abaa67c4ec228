N=$1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.log 2> gpurun_out/r2_bench_n$N.err
tail -c 400 gpurun_out/r2_bench_n$N.err
python scripts/_show_oc.py gpurun_out/r2_bench_n$N.log 2500 | head -4
