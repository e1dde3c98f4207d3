#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --log-n 18"
for mc in 4 8 16 32 64; do
  PB_LOGUP_JIT_CHUNKS=$mc timeout 600 $B > gpurun_out/tune_mc$mc.json 2> gpurun_out/tune_mc$mc.err
  python - "$mc" gpurun_out/tune_mc$mc.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); s=d['stages_ms']
    print('module chunks', sys.argv[1], 'lde %.2f logup_gen %.2f quotient %.2f logup_commit %.2f total %.2f keygen %.1f' % (s['lde'], s['logup_gen'], s['quotient'], s['logup_commit'], s['total'], d['keygen_s']))
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
PB_LDE_NO_TMA=1 timeout 600 $B > gpurun_out/tune_notma.json 2> gpurun_out/tune_notma.err; python -c "
import json; d=json.load(open('gpurun_out/tune_notma.json')); s=d['stages_ms']; print('no-TMA: lde %.2f logup_commit %.2f total %.2f' % (s['lde'], s['logup_commit'], s['total']))"
