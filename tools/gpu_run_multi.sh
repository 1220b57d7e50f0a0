#!/bin/bash
# 2-GPU validation: bench under torchrun (NCCL), reference arm, sharded_ntt with NCCL all-gather
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py > gpurun_out/dist_check.log 2>&1
cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err; tail -5 gpurun_out/dist_check.log
