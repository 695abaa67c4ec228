set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
$B > gpurun_out/r2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_l.log 2>&1
B4="python bench.py --batch 444 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
$B4 > gpurun_out/r2_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:svf_grid5 -s 3 -c 1 -f -o gpurun_out/r2_svf_grid5 $B4 > gpurun_out/r2_ncu_g5.log 2>&1
python scripts/prof_dense.py > gpurun_out/r2_plain_dense.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_gemm -s 8 -c 1 -f -o gpurun_out/r2_dense_gemm python scripts/prof_dense.py > gpurun_out/r2_ncu_dense.log 2>&1
python scripts/prof_slab_flow.py 128 > gpurun_out/r2_plain_flow.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:slab_flow -s 1 -c 1 -f -o gpurun_out/r2_slab_flow python scripts/prof_slab_flow.py 128 > gpurun_out/r2_ncu_flow.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
tail -3 gpurun_out/r2_ncu_flow.log gpurun_out/r2_ncu_dense.log gpurun_out/r2_ncu_g5.log
