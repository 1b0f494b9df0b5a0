#!/bin/bash
python -m pytest tests/test_gpu_motion.py tests/test_gpu_api.py -x -q 2>&1 | tail -3
FSG_FWD_WARP_MIN_TAPS=1 python -m pytest tests/test_gpu_motion.py -x -q 2>&1 | tail -2
python tools/bench_configs.py --config motion --reps 5 2>/dev/null > gpurun_out/motion_r1l.jsonl; python -c "
import sys, json
for l in open('gpurun_out/motion_r1l.jsonl'):
    d=json.loads(l); print(d['psf_taps'], 'fwd %.2f xpairs %.2f xyquads %.2f | ref-ext %.2f' % (d['forward_ms_ours'], d['forward_ms_ours_xpairs'], d['forward_ms_ours_xyquads'], d.get('forward_ms_reference_ext',0)))"
python tools/bench_configs.py --config artifacts --reps 5 > gpurun_out/artifacts6.json 2>gpurun_out/artifacts6.err
python -c "
import json;d=json.load(open('gpurun_out/artifacts6.json'));print({k:(round(v['ms_mean'],2), round(v['ms_min'],1), round(v['ms_max'],1)) for k,v in d['artifacts_ms'].items()}, d['volumes_per_s_single_stream'])"
