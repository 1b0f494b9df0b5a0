#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_motion.py tests/test_gpu_artifacts.py tests/test_gpu_api.py -x -q 2>&1 | tail -3
FSG_FWD_WARP_MIN_TAPS=1 python -m pytest tests/test_gpu_motion.py -x -q 2>&1 | tail -2
python scratch/motionprof.py 2>&1 | grep "^wall" | cut -c1-150
python tools/bench_configs.py --config artifacts --reps 5 > gpurun_out/artifacts5.json 2>gpurun_out/artifacts5.err
python -c "
import json;d=json.load(open('gpurun_out/artifacts5.json'));print({k:(round(v['ms_mean'],2), round(v['ms_min'],1), round(v['ms_max'],1)) for k,v in d['artifacts_ms'].items()}, d['volumes_per_s_single_stream'])"
