#!/bin/bash
# Round-2 measurement pass on one B200: the bench line (all BASELINE configs as extra_workloads), the reference arm, the
# front-end micro-benchmarks, the GRU kernel sweep, the ncu launch list of the default bench command and `--set full`
# captures of the kernels.  Everything lands in gpurun_out/; scripts/profile_summary.py / profile_hotlines.py turn the
# captures into profiles/r02_*.txt (run where ncu is installed).
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
timeout 120 python scripts/bench_fbank.py > gpurun_out/fbank_r02.json 2> gpurun_out/fbank_r02.err
timeout 120 python scripts/bench_mfcc.py > gpurun_out/mfcc_r02.json 2> gpurun_out/mfcc_r02.err
timeout 200 python scripts/gru_sweep.py > gpurun_out/gru_sweep_r02.jsonl 2> gpurun_out/gru_sweep_r02.err
timeout 200 python scripts/gru_sweep.py small > gpurun_out/gru_sweep_small_r02.jsonl 2>> gpurun_out/gru_sweep_r02.err
timeout 60 build/mma_rate_probe > gpurun_out/mma_rate_r02.jsonl 2>&1
timeout 60 build/l2_stream_probe > gpurun_out/l2_stream_r02.jsonl 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches_r02.log 2>&1
cap() {   # name, kernel regex, command...
  local name=$1 k=$2; shift 2
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o gpurun_out/prof_r02_$name "$@" > gpurun_out/prof_r02_$name.log 2>&1
}
cap mdtc mdtc_tc_kernel python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras
cap fbank fbank_kernel python bench.py --workload pcm_e2e_1250x1s --steps 10 --warmup 3 --no-cpu-baseline --no-extras
cap tcn tcn_tc_kernel python bench.py --workload tcn_b1024_t40 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
cap dstcn dstcn_tc_kernel python bench.py --workload ds_tcn_b1024_t40 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
cap gru_fp32 gru_kernel python bench.py --workload gru_b512_t1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
cap gru_tc gru_tc_kernel python scripts/gru_sweep.py
tail -c 400 gpurun_out/bench_r02.json
