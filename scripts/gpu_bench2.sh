#!/bin/bash
# full-size parity + N=1 bench (both arms) + TMA on/off at 2^20 + launch list + ncu of the dominant kernel (summarised on the box)
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled > gpurun_out/thp.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_fullsize.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fullsize.txt
tail -4 gpurun_out/pytest_fullsize.txt
timeout 1500 python bench.py --steps 5 --warmup 3 --metrics-out gpurun_out/metrics_v1.json > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
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
PB_LDE_NO_TMA=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_notma.json 2> gpurun_out/bench_notma.err
python -c "
import json; d=json.load(open('gpurun_out/bench_notma.json')); s=d['stages_ms']; print('no-TMA: lde %.2f logup_commit %.2f total %.2f' % (s['lde'], s['logup_commit'], s['total']))"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-1800 gpurun_out/bench_ref.json
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
for k in leaf_hash_cols transposed_tma; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o /tmp/$k -f $BENCH > gpurun_out/ncu_$k.log 2>&1
  ncu -i /tmp/$k.ncu-rep --page raw --csv > gpurun_out/ncu_$k.raw.csv 2>/dev/null
done
du -sh gpurun_out
