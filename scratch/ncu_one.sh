#!/bin/bash
# usage: scratch/ncu_one.sh TAG KERNEL_REGEX [skip] [count]
TAG=$1; K=$2; S=${3:-3}; C=${4:-1}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
