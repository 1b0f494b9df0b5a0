#!/bin/bash
# usage: scratch/gpu_round.sh TAG  — tests, bench, launch list, full ncu capture of one step
TAG=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/benchref_$TAG.json 2> gpurun_out/benchref_$TAG.err; echo "ref rc=$?"; cat gpurun_out/benchref_$TAG.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "launches rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gmm_kernel|warp_fast|warp_kernel|sep_|zoom_rows" -s ${NCU_SKIP:-27} -c ${NCU_COUNT:-9} -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
