timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_final.log 2> gpurun_out/r2_bench_n8_final.err
tail -c 300 gpurun_out/r2_bench_n8_final.err
python scripts/show_bench_line.py gpurun_out/r2_bench_n8_final.log 1100 | head -3
