#!/bin/bash
# one gpurun call: micro-benchmarks, GPU parity tests, smoke, first bench lines.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( timeout 120 ./scripts/microbench > gpurun_out/microbench.txt 2>&1 ) || echo "microbench failed"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 600 python bench.py --log-n 16 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_2p16.json 2> gpurun_out/bench_2p16.err; tail -c 1500 gpurun_out/bench_2p16.json; tail -3 gpurun_out/bench_2p16.err
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
