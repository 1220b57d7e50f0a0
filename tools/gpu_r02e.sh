#!/bin/bash
# multi-GPU diagnosis: which step of the assembly modes faults (CUDA_LAUNCH_BLOCKING localises it)
N=${1:-2}
mkdir -p gpurun_out
for cfg in "nccl 1" "nccl 0" "peers 1" "peers 0"; do
  set -- $cfg
  echo "=== stage $1 SA_NTT_PDL=$2" >> gpurun_out/r02e_dist_debug.log
  CUDA_LAUNCH_BLOCKING=1 SA_NTT_PDL=$2 timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dist_debug.py $1 2>&1 | grep -E "^\[rank|Error|error|rror:" | head -40 >> gpurun_out/r02e_dist_debug.log
done
cat gpurun_out/r02e_dist_debug.log | cut -c1-300
