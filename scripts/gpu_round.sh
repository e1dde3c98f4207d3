#!/bin/bash
# one gpurun call: GPU parity tests, smoke, bench lines (+ optional profile).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -6 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 900 python bench.py --steps 3 --warmup 2 ${BENCH_EXTRA} > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_full.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
    print('stages',d['stages_ms']); print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['stage_roofline_frac'])
    print('e2e',d.get('e2e',{}).get('value'), d.get('e2e',{}).get('stages_ms'))
    print('cpu',d.get('cpu_baseline',{}).get('value'))
except Exception as e: print('bench parse failed',e)
PY
tail -5 gpurun_out/bench_full.err
if [ -n "$PROFILE" ]; then bash scripts/gpu_profile.sh; fi
