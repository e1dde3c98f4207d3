#!/bin/bash
# 2 GPUs, NCCL: the sharded prover + its query phase through TorchComm at a small and a mid size (proof and query openings compared with
# the single-GPU ones inside bench.py: proof_equals_single_gpu)
mkdir -p gpurun_out
for cfg in "14 64 20 40" "17 256 40 200"; do
set -- $cfg
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --log-n $1 --width $2 --constraints $3 --interactions $4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_2gpu_$1.json 2> gpurun_out/bench_2gpu_$1.err
echo "exit $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_2gpu_$1.json') if l.startswith('{')][-1])
    print('2^$1', d['value'], d.get('proof_equals_single_gpu'), d.get('speedup'), {k:d[k] for k in d if 'e2e' in k or 'replica' in k})
except Exception as e: print('parse failed', e)
PY
tail -3 gpurun_out/bench_2gpu_$1.err
done
