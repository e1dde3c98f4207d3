#!/bin/bash
# multi-GPU bench lines: $1 = N, $2 = workload (keccak|multichip|pairing), rest = extra args
N=$1; WL=$2; shift 2
mkdir -p gpurun_out
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL "$@" > gpurun_out/bench_${WL}_${N}gpu.json 2> gpurun_out/bench_${WL}_${N}gpu.err
echo "exit $?"; tail -c 3000 gpurun_out/bench_${WL}_${N}gpu.json; tail -5 gpurun_out/bench_${WL}_${N}gpu.err
