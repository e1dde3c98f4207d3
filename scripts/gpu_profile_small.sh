#!/bin/bash
# refresh of the launch list and of the opening / quotient kernel captures (after their rewrite)
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_d.csv $BENCH > gpurun_out/launches_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:deep_quotient_kernel|eval_partial_kernel|pbq" -s 7 -c 7 -o gpurun_out/open_d -f $BENCH > gpurun_out/ncu_open_d.log 2>&1
ls -la gpurun_out | tail -6
