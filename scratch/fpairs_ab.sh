#!/bin/bash
python -m pytest tests/test_gpu_base.py -x -q -k "pairs" 2>&1 | tail -6
for mode in 0 2 1; do
FSG_GMM_PAIRS=$mode python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d['roofline']['per_call_ms']; print('mode $mode', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'gmm', p['fsg_gmm'], 'warp', p['fsg_warp'], 'frac', round(d['roofline']['frac'],3))
"
done
