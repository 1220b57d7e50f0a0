#!/bin/bash
# 8-GPU box: the scaling line of the bench at N = 8, 4, 2 and the sharded-NTT / Merkle NCCL check
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${n}gpu.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 tools/dist_check.py > gpurun_out/dist_check8.log 2>&1; tail -3 gpurun_out/dist_check8.log
