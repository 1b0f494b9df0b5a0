#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bench_configs.py --config motion --reps 1"
$CMD > gpurun_out/motion_plain.log 2>&1; echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"slice_adj|slice_fwd|equalize" -c 12 -f -o gpurun_out/prof_motion $CMD > gpurun_out/ncu_motion.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_motion.log
