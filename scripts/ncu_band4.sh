#!/bin/bash
# ncu --set full capture of the five kernels of one two-sided solve (n=1862, bw=320: the C1 normal equations' shape)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES=0 REPS=3 ncu --set full --clock-control none --import-source on -k regex:band_ -s 5 -c 5 -f -o gpurun_out/prof_band4 \
    python scripts/one_band4.py > gpurun_out/ncu_band4.log 2>&1
tail -3 gpurun_out/ncu_band4.log
