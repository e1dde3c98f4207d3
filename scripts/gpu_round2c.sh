#!/bin/bash
# round 2c check: merkle split retune + device-driven FRI commit phase + AVX-512 host transcript + multi-chip prover
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_chips.py tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_tracegen.py -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
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
timeout 900 python bench.py --workload multichip --one-transcript --steps 2 --warmup 1 > gpurun_out/bench_multichip_one.json 2> gpurun_out/bench_multichip_one.err; cut -c1-1500 gpurun_out/bench_multichip_one.json; tail -3 gpurun_out/bench_multichip_one.err
