#!/bin/bash
# Run on the GPU box (through gpurun): ncu captures of the main kernels, one launch each after warm-up.
#   bash profiles/capture.sh <tag> [kernel-regex ...]
# Writes gpurun_out/<tag>_<kernel>.ncu-rep (+ logs). Summarise afterwards with profiles/summarize.py <tag>.
tag=${1:-r02}; shift
kernels=${@:-hist_lane_kernel tables_build_kernel encode_kernel dec_sync_kernel dec_write_kernel}
mkdir -p gpurun_out
for k in $kernels; do
  skip=3   # the launch of the step after three warm-up steps
  [ "$k" = encode_kernel ] && skip=6   # two encoder launches per step (the 64-symbol instance, then the one that leaves at once): the first of the fourth pair
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${tag}_$k \
    python bench.py --config markov --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_${tag}_$k.log 2>&1
  echo "$k: $(grep -c 'Profiling' gpurun_out/ncu_${tag}_$k.log) launch profiled"
done
