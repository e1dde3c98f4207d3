#!/bin/bash
# LogUp JIT knob sweep, second pass (2^18 rows), then ncu captures summarised ON THE BOX (the reports are too big to bring back)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --log-n 18"
for cfg in "1 128 0" "2 128 0" "2 256 0" "2 128 4" "2 256 2" "3 128 0" "1 256 0" "2 64 0" "2 128 6"; do
  set -- $cfg
  PB_LOGUP_GROUP=$1 PB_LOGUP_BLOCK=$2 PB_LOGUP_MINB=$3 timeout 600 $B > gpurun_out/tune_$1_$2_$3.json 2> gpurun_out/tune_$1_$2_$3.err
  python - "$cfg" gpurun_out/tune_$1_$2_$3.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); s=d['stages_ms']
    print('group/block/minb', sys.argv[1], 'logup_gen %.2f quotient %.2f total %.2f keygen %.1f' % (s['logup_gen'], s['quotient'], s['total'], d['keygen_s']))
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
export PB_LOGUP_GROUP=2
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
for k in pbl_perm pbl_fold leaf_hash_cols; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -o /tmp/$k -f $BENCH > gpurun_out/ncu_$k.log 2>&1
  ncu -i /tmp/$k.ncu-rep --page raw --csv > gpurun_out/ncu_$k.raw.csv 2>/dev/null
  ncu -i /tmp/$k.ncu-rep --page source --csv > /tmp/$k.src.csv 2>/dev/null; head -c 3000000 /tmp/$k.src.csv > gpurun_out/ncu_$k.source.csv
done
ls -la gpurun_out | tail -12; du -sh gpurun_out
