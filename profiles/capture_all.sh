#!/bin/bash
# Everything profiles/ holds for one round, in one gpurun call:   bash profiles/capture_all.sh r02
#   1. the default bench line (all configurations), no profiler attached   -> gpurun_out/bench_<tag>.json
#   2. the launch list of a short run (per-kernel share of the step)        -> gpurun_out/launches_<tag>.csv
#   3. DRAM traffic of the main kernels at the bench's full size           -> gpurun_out/traffic_<tag>.csv
#   4. one `ncu --set full` capture per main kernel (bench size)         -> gpurun_out/<tag>_<kernel>.ncu-rep
# Then, back in the container:  python profiles/summarize.py <tag> && python profiles/traffic.py <tag>
tag=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo "bench failed"; tail -5 gpurun_out/bench_$tag.err; exit 1; }
echo "bench: $(cut -c1-160 gpurun_out/bench_$tag.json)"
FULL="python bench.py --config markov --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$FULL > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $FULL > gpurun_out/ncu_list.log 2>&1
echo "launch list: $(grep -c gpu__time_duration gpurun_out/launches_$tag.csv) rows"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'hist_lane_kernel|tables_build_kernel|encode_kernel|dec_sync_kernel|dec_write_kernel' \
    --csv --log-file gpurun_out/traffic_$tag.csv $FULL > gpurun_out/ncu_traffic.log 2>&1
echo "traffic: $(grep -c dram__bytes_read gpurun_out/traffic_$tag.csv) rows"
bash profiles/capture.sh $tag hist_lane_kernel tables_build_kernel encode_kernel dec_sync_kernel dec_write_kernel
