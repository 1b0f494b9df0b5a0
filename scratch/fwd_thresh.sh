#!/bin/bash
for t in 100 10000; do
FSG_FWD_WARP_MIN_TAPS=$t python tools/bench_configs.py --config motion --reps 5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('min_taps $t', d['psf_taps'], 'fwd %.2f xyquads %.2f' % (d['forward_ms_ours'], d['forward_ms_ours_xyquads']))"
done
