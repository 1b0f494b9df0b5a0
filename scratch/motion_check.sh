#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_motion.py tests/test_gpu_artifacts.py tests/test_gpu_api.py -x -q 2>&1 | tail -3
python scratch/motionprof.py 2>&1 | grep "^wall" | cut -c1-120
python tools/bench_configs.py --config artifacts --reps 5 > gpurun_out/artifacts4.json 2>gpurun_out/artifacts4.err
python -c "
import json;d=json.load(open('gpurun_out/artifacts4.json'));print({k:(round(v['ms_mean'],2), round(v['ms_min'],1), round(v['ms_max'],1)) for k,v in d['artifacts_ms'].items()}, d['volumes_per_s_single_stream'])"
python tools/bench_configs.py --config sample_api --reps 5 > gpurun_out/sample_api4.json 2>gpurun_out/sample_api4.err
python -c "
import json;d=json.load(open('gpurun_out/sample_api4.json'))
for k,v in d['results'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('median_s','mean_s','min_s','max_s','cold_start_first_call_s')})"
