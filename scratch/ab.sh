#!/bin/bash
# A/B timing of kernel variants: prints per-call ms of each entry point for each library
python -m pytest tests/test_gpu_base.py tests/test_host.py -x -q 2>&1 | tail -3
for lib in fetalsyngen_b200/libfsg.so "$@"; do
  echo "== $lib"
  FSG_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${AB_ARGS:---no-e2e} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['per_call_ms'])
    else: print(l.rstrip()[-300:])
"
done
