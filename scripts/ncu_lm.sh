#!/bin/bash
# ncu launch list of one bench command + full capture of the J^T J kernel (B200_PROFILING.md recipe)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:data_jtj_kernel -s 30 -c 3 -f -o gpurun_out/prof_jtj \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_jtj.log 2>&1
ls -la gpurun_out
