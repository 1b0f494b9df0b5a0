#!/bin/bash
python -m pytest tests/test_gpu_api.py -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_packed.json 2> gpurun_out/bench_packed.err; echo rc=$?; tail -3 gpurun_out/bench_packed.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_packed.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e'], d['e2e_packed_inputs'])"
