#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_base.py tests/test_gpu_api.py tests/test_gpu_fullsize.py tests/test_gpu_seeds.py -x -q > gpurun_out/pytest_zoom.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_zoom.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_zoomA.json 2>gpurun_out/bench_zoomA.err; echo rc=$?; python -c "
import json;d=json.loads(open('gpurun_out/bench_zoomA.json').read().strip().splitlines()[-1]);print('pruned',d['value'],d['ms_per_step'],d['roofline']['per_call_ms'])"
FSG_ZOOM_FULL_REDUCE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_zoomB.json 2>gpurun_out/bench_zoomB.err; echo rc=$?; python -c "
import json;d=json.loads(open('gpurun_out/bench_zoomB.json').read().strip().splitlines()[-1]);print('full  ',d['value'],d['ms_per_step'],d['roofline']['per_call_ms'])"
