#!/bin/bash
# One profiling round on the GPU box (run under gpurun): usage tools/profile_round.sh TAG
#   1. the bench command runs plain first (a number printed under ncu is never a bench value);
#   2. launch list: gpu__time_duration.sum of every kernel of two steps;
#   3. --set full capture of the step's main kernels (one launch each) with source counters.
# Outputs land in gpurun_out/<TAG>_*; summarise here with tools/ncu_summary.py / tools/ncu_hot.py.
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gmm_kernel|warp_fast_kernel|sep_stream_kernel|sep_zrow_kernel|zoom_walk_kernel" -s 21 -c 7 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu_full.log
