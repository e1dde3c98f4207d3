#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded.txt
tail -15 gpurun_out/pytest_sharded.txt
