#!/bin/bash
# Round-1 measurement pass on one B200: bench lines for every workload, the reference arm, the fbank micro-bench,
# the ncu launch list of the default bench command and `--set full` captures of the tensor-core kernels.
# Everything lands in gpurun_out/; scripts/profile_summary.py turns the captures into profiles/r01_*.txt.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
for w in gru_b512_t1 tcn_b1024_t40 ds_tcn_b1024_t40 pcm_e2e_1250x1s; do
  timeout 300 python bench.py --workload $w > gpurun_out/bench_r01_$w.json 2> gpurun_out/bench_r01_$w.err
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err
timeout 120 python scripts/bench_fbank.py > gpurun_out/fbank_r01.json 2> gpurun_out/fbank_r01.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r01.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mdtc_tc_kernel -s 5 -c 1 -f -o gpurun_out/prof_mdtc_tc_r01 \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/prof_mdtc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tcn_tc_kernel -s 5 -c 1 -f -o gpurun_out/prof_tcn_tc_r01 \
  python bench.py --workload tcn_b1024_t40 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/prof_tcn.log 2>&1
tail -c 600 gpurun_out/bench_r01.json
