#!/bin/bash
# round 2d check: 16-lane permutations for the latency-bound levels (tree tops, small FRI layers, transcript kernel)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_chips.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('keccak', d['value'], d['stages_ms'])"
tail -2 gpurun_out/bench_quick.err
for ln in 12 16; do
timeout 300 python bench.py --log-n $ln --width 128 --constraints 70 --interactions 110 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_small_$ln.json 2> gpurun_out/bench_small_$ln.err
python -c "
import json; d=json.load(open('gpurun_out/bench_small_$ln.json')); print('small 2^$ln x128', d['value'], d['stages_ms'])"
done
