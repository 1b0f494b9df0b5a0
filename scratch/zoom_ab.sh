#!/bin/bash
for tag in nr1 nr2mb2 nr2mb3 nr1mb4; do
FSG_LIB=$PWD/scratch/libfsg_$tag.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d['roofline']['per_call_ms']; print('$tag', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'minmax', p['fsg_zoom_minmax'], 'zoom', p['fsg_zoom'])
"
done
