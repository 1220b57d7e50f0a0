#!/bin/bash
# 8-GPU sweep of the push-kernel variants (SA_PUSH_MODE 0 = rotate over peers, 1 = one peer per CTA, 2 = TMA bulk
# stores) next to the fused peer-store kernel, then the bench line
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for cfg in "0 148" "1 147" "1 294" "2 148" "2 74" "2 296"; do
  set -- $cfg; i=$((i+1))
  SA_DIST_SKIP_FRI=1 SA_DIST_MODES=p2p-push,p2p-store SA_PUSH_MODE=$1 SA_PUSH_CTAS=$2 timeout 240 $TR --master-port $((29520+i)) tools/dist_check.py > gpurun_out/r02m_dist_check_${N}gpu_mode$1_ctas$2.log 2>&1
done
timeout 700 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/r02m_bench_${N}gpu.json 2> gpurun_out/r02m_bench_${N}gpu.err
for f in gpurun_out/r02m_dist_check_${N}gpu_*.log; do echo $f; grep DIST_CHECK $f | head -1 | cut -c1-500; grep -i "error" $f | head -2 | cut -c1-200; done; python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_${N}gpu.json')); print(d['value'], d['e2e']['value'], json.dumps(d['with_allgather'])[:1200])"
