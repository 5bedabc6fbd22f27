#!/bin/bash
# ncu --set full of ONE kernel of the trunk (regex $1), source-level counters exported as CSV; .ncu-rep stays on the box.
# usage: tools/ncu_one_kernel.sh <kernel-regex> <tag> [launch-skip]
set -x
K=${1:-conv_stem7p}; TAG=${2:-stem}; SKIP=${3:-2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o /tmp/one_${TAG} -f \
    python tools/profile_layers.py 256 bf16 resnet50 > gpurun_out/ncu_one_${TAG}.log 2>&1
ncu -i /tmp/one_${TAG}.ncu-rep --page raw --csv > gpurun_out/one_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/one_${TAG}.ncu-rep --page source --csv --print-source cuda > gpurun_out/one_${TAG}_source.csv 2>/dev/null
ncu -i /tmp/one_${TAG}.ncu-rep --page details > gpurun_out/one_${TAG}_details.txt 2>/dev/null
ls -la gpurun_out/one_${TAG}*; du -sh gpurun_out
