#!/bin/bash
# one gpurun call: full-size sampled parity + the N=1 bench line (both arms); outputs land in gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_fullsize.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fullsize.txt
tail -15 gpurun_out/pytest_fullsize.txt
timeout 1500 python bench.py --steps ${STEPS:-5} --warmup ${WARMUP:-3} ${BENCH_EXTRA} > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_full.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks','keygen_s','perm_width')})
    print('stages',d['stages_ms']); print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['stage_roofline_frac'])
    print('e2e',d.get('e2e',{}).get('value'), d.get('e2e',{}).get('stages_ms'))
    print('cpu',d.get('cpu_baseline'))
except Exception as e: print('bench parse failed',e)
PY
tail -5 gpurun_out/bench_full.err
if [ -n "$REFARM" ]; then timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-1500; fi
