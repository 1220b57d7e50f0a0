#!/bin/bash
# multi-GPU pass (gpurun --gpus N): every assembly mode of sharded_ntt over N ranks + sharded FRI instances
# (tools/dist_check.py), variants of the peer modes, then the bench line with its with_allgather leg
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02g_topo_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r02g_dist_check_${N}gpu.log 2>&1
SA_DIST_SKIP_FRI=1 SA_DIST_MODES=p2p-store,p2p-push SA_NTT_PEER_C=8 SA_PUSH_CTAS=64 timeout 300 $TR --master-port 29513 tools/dist_check.py > gpurun_out/r02g_dist_check_${N}gpu_v1.log 2>&1
SA_DIST_SKIP_FRI=1 SA_DIST_MODES=p2p-push SA_PUSH_CTAS=296 timeout 300 $TR --master-port 29514 tools/dist_check.py > gpurun_out/r02g_dist_check_${N}gpu_v2.log 2>&1
SA_DIST_SKIP_FRI=1 SA_DIST_MODES=p2p-push SA_PUSH_CTAS=32 timeout 300 $TR --master-port 29515 tools/dist_check.py > gpurun_out/r02g_dist_check_${N}gpu_v3.log 2>&1
timeout 700 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/r02g_bench_${N}gpu.json 2> gpurun_out/r02g_bench_${N}gpu.err
for f in gpurun_out/r02g_dist_check_${N}gpu*.log; do grep DIST_CHECK $f | head -1 | cut -c1-900; done; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_${N}gpu.json')); print(d['value'], d['e2e']['value'], json.dumps(d['with_allgather'])[:1500])"; tail -3 gpurun_out/r02g_bench_${N}gpu.err | cut -c1-300
