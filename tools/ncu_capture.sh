#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time of exactly one bench step (eager, no graph),
# (2) one --set full capture of every conv-family launch (conv_*, conv3x3_halo, bottleneck64) of ONE STEP = both trunks (the .ncu-rep stays on the box; the raw and
# source-counter CSV pages come back).  1 GPU only.  Numbers under ncu are never bench values.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-range > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:conv|bottleneck' -c 120 -o /tmp/conv_${TAG} -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-range > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i /tmp/conv_${TAG}.ncu-rep --page raw --csv > gpurun_out/conv_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/conv_${TAG}.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|Throughput|Tensor|Registers|Warp Cycles|Theoretical Occ|Achieved Occ|Shared Memory Config|Block Limit" | head -400 > gpurun_out/conv_${TAG}_details.csv
cuobjdump -sass spec_b200/libspecb200.so 2>/dev/null | grep -oE "\b(UTC[A-Z]*MMA[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|HMMA[A-Z0-9_.]*)" | sort | uniq -c > gpurun_out/sass_mnemonics_${TAG}.txt
ls -la gpurun_out/ | tail -8; du -sh gpurun_out
