#!/bin/bash
# NVLS multicast assembly (multimem stores through one multicast address) next to the fused peer-store kernel,
# N GPUs (first argument)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SA_DIST_SKIP_FRI=1 SA_DIST_MODES=p2p-store,nvls-store,nvls-push,p2p-push timeout 300 $TR --master-port 29541 tools/dist_check.py > gpurun_out/r02n_dist_check_${N}gpu.log 2>&1
grep DIST_CHECK gpurun_out/r02n_dist_check_${N}gpu.log | head -1 | cut -c1-900
grep -i "error\|Traceback" gpurun_out/r02n_dist_check_${N}gpu.log | head -8 | cut -c1-300
