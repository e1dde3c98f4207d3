#!/bin/bash
# ncu evidence for one round: launch list of one full-size step + full captures of the dominant kernels.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:leaf_hash_cols -s 2 -c 1 -o gpurun_out/leaf -f $BENCH > gpurun_out/ncu_leaf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nttf -s 440 -c 4 -o gpurun_out/ntt -f $BENCH > gpurun_out/ncu_ntt.log 2>&1
ls -la gpurun_out | tail -12
