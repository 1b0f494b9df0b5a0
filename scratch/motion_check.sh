#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_motion.py tests/test_gpu_artifacts.py -x -q 2>&1 | tail -3
python scratch/motionprof.py 2>&1 | grep "^wall"
python tools/bench_configs.py --config artifacts --reps 5 > gpurun_out/artifacts3.json 2>gpurun_out/artifacts3.err
python -c "
import json;d=json.load(open('gpurun_out/artifacts3.json'));print({k:round(v['ms_mean'],2) for k,v in d['artifacts_ms'].items()}, d['volumes_per_s_single_stream'])"
