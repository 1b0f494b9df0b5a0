#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), d['roofline']['per_call_ms'])
    else: print(l.rstrip()[-300:])
"
