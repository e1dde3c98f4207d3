#!/bin/bash
# last call of the round: the final tree's GPU tests (everything but the two full-size tests) + smoke
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_chips.py tests/test_gpu_sharded.py tests/test_gpu_parity.py tests/test_gpu_tracegen.py -m gpu -q -x > gpurun_out/pytest_last.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_last.txt
tail -5 gpurun_out/pytest_last.txt
timeout 60 python __graft_entry__.py --smoke > gpurun_out/smoke_last.txt 2>&1; tail -2 gpurun_out/smoke_last.txt
