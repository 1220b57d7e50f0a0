#!/bin/bash
# final-code check of the torchrun bench line at N GPUs (first argument): value, e2e, with_allgather modes
N=${1:-2}
mkdir -p gpurun_out
SA_BENCH_NVLS=1 timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/r02z_bench_${N}gpu.json 2> gpurun_out/r02z_bench_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_${N}gpu.json')); print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print(json.dumps(d['with_allgather'])[:1500])"
tail -3 gpurun_out/r02z_bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02z_bench_ref_${N}gpu.json 2>> gpurun_out/r02z_bench_${N}gpu.err; cut -c1-300 gpurun_out/r02z_bench_ref_${N}gpu.json
