#!/bin/bash
bash scratch/gpu_round.sh r1j
python tools/bench_configs.py --config motion --reps 3 > gpurun_out/motion_r1j.jsonl 2> gpurun_out/motion_r1j.err; echo "motion rc=$?"; head -c 600 gpurun_out/motion_r1j.jsonl; echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
