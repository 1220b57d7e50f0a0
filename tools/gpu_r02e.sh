#!/bin/bash
# multi-GPU diagnosis: which step of the assembly modes faults (CUDA_LAUNCH_BLOCKING localises it)
N=${1:-2}
mkdir -p gpurun_out
for cfg in "peers 1 16" "peers 1 20" "nccl 1 20"; do
  set -- $cfg
  echo "=== stage $1 SA_NTT_PDL=$2 log_n=$3" >> gpurun_out/r02e_dist_debug.log
  CUDA_LAUNCH_BLOCKING=1 SA_NTT_PDL=$2 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dist_debug.py $1 $3 2>&1 | grep -E "^\[rank|Error|error|rror:" | head -40 >> gpurun_out/r02e_dist_debug.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r02e_dist_check_${N}gpu.log 2>&1
cat gpurun_out/r02e_dist_debug.log | cut -c1-300; grep DIST_CHECK gpurun_out/r02e_dist_check_${N}gpu.log | cut -c1-1200
