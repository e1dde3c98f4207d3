#!/bin/bash
# ncu evidence for one round: launch list of one full-size step + full captures of the dominant kernels.
# (kernel-name filters are matched against the UNQUALIFIED function name)
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:leaf_hash_cols -s 2 -c 1 -o gpurun_out/leaf -f $BENCH > gpurun_out/ncu_leaf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:strided_kernel|transposed_kernel" -s 8 -c 4 -o gpurun_out/ntt -f $BENCH > gpurun_out/ncu_ntt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:deep_quotient_kernel|eval_partial_kernel|pbq" -s 3 -c 3 -o gpurun_out/open -f $BENCH > gpurun_out/ncu_open.log 2>&1
ls -la gpurun_out | tail -12
