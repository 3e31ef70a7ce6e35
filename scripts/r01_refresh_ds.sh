set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -3
timeout 300 python bench.py --workload ds_tcn_b1024_t40 > gpurun_out/bench_r01_ds_tcn_b1024_t40.json 2> gpurun_out/bench_r01_ds_tcn_b1024_t40.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dstcn_tc_kernel -s 5 -c 1 -f -o gpurun_out/prof_dstcn_tc_r01 python bench.py --workload ds_tcn_b1024_t40 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/prof_dstcn.log 2>&1
tail -c 900 gpurun_out/bench_r01_ds_tcn_b1024_t40.json
