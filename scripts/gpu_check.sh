#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tracegen.py tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -6 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --workload stage0 --steps 6 > gpurun_out/bench_stage0.json 2> gpurun_out/bench_stage0.err; cat gpurun_out/bench_stage0.json; tail -3 gpurun_out/bench_stage0.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['stages_ms'])"
timeout 900 python bench.py --workload multichip --steps 2 --warmup 1 > gpurun_out/bench_multichip_1gpu.json 2> gpurun_out/bench_multichip_1gpu.err; cut -c1-600 gpurun_out/bench_multichip_1gpu.json
