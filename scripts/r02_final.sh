#!/bin/bash
# Final round-2 measurement pass on one B200 (bounded: ~9 minutes): the GPU parity suite, the bench line (all BASELINE
# configs as extra_workloads), the front-end micro-benchmarks, the ncu launch list of the default bench command and
# `--set full` captures of the kernels that changed since the last pass (fbank, tcn).  Everything lands in gpurun_out/;
# scripts/profile_summary.py / profile_hotlines.py turn the captures into profiles/r02_*.txt (run where ncu is installed).
set -u
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/final_tests.txt
timeout 400 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
timeout 100 python scripts/bench_fbank.py > gpurun_out/fbank_r02.json 2> gpurun_out/fbank_r02.err
timeout 100 python scripts/bench_mfcc.py > gpurun_out/mfcc_r02.json 2> gpurun_out/mfcc_r02.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches_r02.log 2>&1
cap() {   # name, kernel regex, command...
  local name=$1 k=$2; shift 2
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o gpurun_out/prof_r02_$name "$@" > gpurun_out/prof_r02_$name.log 2>&1
}
cap fbank fbank_kernel python bench.py --workload pcm_e2e_1250x1s --steps 10 --warmup 3 --no-cpu-baseline --no-extras
cap tcn tcn_tc_kernel python bench.py --workload tcn_b1024_t40 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1
cat gpurun_out/final_tests.txt; tail -2 gpurun_out/final_smoke.txt; tail -c 300 gpurun_out/bench_r02.json
