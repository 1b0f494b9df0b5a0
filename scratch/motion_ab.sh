#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_motion.py -x -q > gpurun_out/pytest_motion.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_motion.log
python tools/bench_configs.py --config motion --reps 5 > gpurun_out/motion_lean.jsonl 2>gpurun_out/motion_lean.err; echo rc=$?
FSG_ADJ_LEAN=0 FSG_FWD_LEAN=0 python tools/bench_configs.py --config motion --reps 5 > gpurun_out/motion_ref.jsonl 2>gpurun_out/motion_ref.err; echo rc=$?
python - <<'PY'
import json
for tag in ("lean","ref"):
    for l in open(f"gpurun_out/motion_{tag}.jsonl"):
        d=json.loads(l); print(tag, d["psf_taps"], "fwd %.2f adj %.2f  ref-ext fwd %.2f adj %.2f" % (d["forward_ms_ours"], d["adjoint_ms_ours"], d.get("forward_ms_reference_ext",0), d.get("adjoint_ms_reference_ext",0)))
PY
python tools/bench_configs.py --config artifacts --reps 5 > gpurun_out/artifacts2.json 2>gpurun_out/artifacts2.err; python -c "
import json;d=json.load(open('gpurun_out/artifacts2.json'));print({k:round(v['ms_mean'],2) for k,v in d['artifacts_ms'].items()}, d['volumes_per_s_single_stream'])"
