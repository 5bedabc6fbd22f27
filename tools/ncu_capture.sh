#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time of exactly one bench step (eager, no graph),
# (2) one --set full capture of the conv kernels of one step.  1 GPU only.  Numbers under ncu are never bench values.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-range > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -c 110 -o gpurun_out/conv_tc_${TAG} -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-range > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/conv_tc_${TAG}.ncu-rep --page raw --csv > gpurun_out/conv_tc_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
