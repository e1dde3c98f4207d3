#!/bin/bash
# one gpurun call: GPU parity tests + smoke; outputs land in gpurun_out/
mkdir -p gpurun_out
(nproc; free -g | head -2; lscpu | grep -E "Model name|Flags" | cut -c1-300) > gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
