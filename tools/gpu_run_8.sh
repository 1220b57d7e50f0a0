#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
tail -c 600 gpurun_out/bench_${n}gpu.json; echo
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_8gpu.json 2> gpurun_out/bench_ref_8gpu.err
tail -c 300 gpurun_out/bench_ref_8gpu.json
