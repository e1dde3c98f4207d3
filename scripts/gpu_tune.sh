#!/bin/bash
# one gpurun call: Poseidon2 build variants, LogUp JIT knob sweep (2^18 rows), launch list + ncu captures at full size
mkdir -p gpurun_out
python scripts/p2_variants.py > gpurun_out/p2_variants.txt 2>&1; cat gpurun_out/p2_variants.txt
python scripts/leaf_variants.py > gpurun_out/leaf_variants.txt 2>&1; cat gpurun_out/leaf_variants.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --log-n 18"
for cfg in "8 128 0" "4 128 0" "8 256 0" "4 256 2" "8 64 0" "2 128 0" "4 128 4" "8 128 3"; do
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
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbl_perm -s 16 -c 1 -o gpurun_out/lu_perm -f $BENCH > gpurun_out/ncu_lu_perm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbl_fold -s 16 -c 1 -o gpurun_out/lu_fold -f $BENCH > gpurun_out/ncu_lu_fold.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:leaf_hash_cols -s 3 -c 1 -o gpurun_out/leaf -f $BENCH > gpurun_out/ncu_leaf.log 2>&1
ls -la gpurun_out | tail -8
