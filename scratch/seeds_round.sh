#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_seeds.py -x -q > gpurun_out/pytest_seeds.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_seeds.log
python tools/bench_configs.py --config seeds --reps 3 > gpurun_out/seeds_bench.json 2> gpurun_out/seeds_bench.err; echo "bench rc=$?"; cat gpurun_out/seeds_bench.json; tail -5 gpurun_out/seeds_bench.err
