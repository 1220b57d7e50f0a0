#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/e2e_sweep.txt
for cfg in "3 16" "4 16" "2 16" "3 32" "4 32" "6 16" "4 8" "3 64"; do
  set -- $cfg
  SA_HOST_STREAMS=$1 SA_HOST_CHUNK_MIB=$2 timeout 300 python bench.py --steps 50 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('streams $1 chunk $2 MiB e2e', d['e2e']['value'])" >> gpurun_out/e2e_sweep.txt
done
cat gpurun_out/e2e_sweep.txt
