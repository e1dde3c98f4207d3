#!/bin/bash
# round-end evidence in one call: all GPU tests + smoke + N=1 bench (both arms) + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -6 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 1500 python bench.py --steps 5 --warmup 3 --metrics-out gpurun_out/metrics_v1.json > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_full.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks','keygen_s','perm_width')})
    print('stages',d['stages_ms']); print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['stage_roofline_frac'])
    print('e2e',d.get('e2e',{}).get('value'), d.get('e2e',{}).get('stages_ms'))
    print('cpu',d.get('cpu_baseline'))
except Exception as e: print('bench parse failed',e)
PY
tail -3 gpurun_out/bench_full.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-700 gpurun_out/bench_ref.json
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
wc -l gpurun_out/launches.csv
du -sh gpurun_out
