#!/bin/bash
# compute-sanitizer over the hand-written kernels: memcheck, racecheck and synccheck of the GPU parity
# subset that touches every kernel family once (conv kernels incl. the CTA-pair / halo / bottleneck / stem variants, the tail,
# the golden whole-path case).  Small shapes only: the sanitizer slows kernels 10-100x.  Logs land in gpurun_out/ and the
# summaries are copied to profiles/ by hand.  1 GPU.  usage: tools/sanitize.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
SEL='conv_kernels and bf16 or fused_bottleneck and bf16 and not h56 or stem_conv7x7 or against_committed_golden'
for tool in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
        python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_${tool}_${TAG}.log 2>&1
    echo "rc=$?" >> gpurun_out/sanitizer_${tool}_${TAG}.log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitizer_${tool}_${TAG}.log | tail -4
done
