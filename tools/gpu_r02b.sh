#!/bin/bash
# round 2, multi-GPU pass (run with gpurun --gpus N): sa_ntt_multi on one device, every assembly mode of
# sharded_ntt over N ranks, sharded FRI instances, then the bench line with its with_allgather leg
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02b_topo_${N}gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu.py -q -x -k "ntt_multi or table_cache or pipeline_settings" > gpurun_out/r02b_pytest_${N}gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r02b_dist_check_${N}gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/r02b_bench_${N}gpu.json 2> gpurun_out/r02b_bench_${N}gpu.err
tail -3 gpurun_out/r02b_pytest_${N}gpu.log; grep DIST_CHECK gpurun_out/r02b_dist_check_${N}gpu.log | cut -c1-900; tail -5 gpurun_out/r02b_dist_check_${N}gpu.log | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_${N}gpu.json')); print(d['value'], d['with_allgather'])"; tail -3 gpurun_out/r02b_bench_${N}gpu.err
