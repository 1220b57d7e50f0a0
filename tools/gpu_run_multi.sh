#!/bin/bash
# N-GPU bench line under torchrun (N = first argument, default 2) + the host-pipeline parity test on GPU 0
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -x -q -k "host_entry" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print($N, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])"
tail -3 gpurun_out/bench_${N}gpu.err
