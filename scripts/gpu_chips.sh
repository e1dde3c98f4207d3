#!/bin/bash
# multi-chip one-transcript prover: parity tests, then the 50-chip segment both ways
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chips.py -m gpu -x -q > gpurun_out/pytest_chips.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_chips.txt
tail -25 gpurun_out/pytest_chips.txt
timeout 900 python bench.py --workload multichip --one-transcript --steps 2 --warmup 1 > gpurun_out/bench_multichip_one.json 2> gpurun_out/bench_multichip_one.err; cut -c1-1500 gpurun_out/bench_multichip_one.json; tail -5 gpurun_out/bench_multichip_one.err
