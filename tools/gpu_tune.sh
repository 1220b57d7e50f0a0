#!/bin/bash
# tile-shape sweep for the NTT kernel: parity at 2^18..2^20 + short bench per shape
mkdir -p gpurun_out
: > gpurun_out/tune.jsonl
for shape in "4 8" "4 4" "4 2" "3 8" "3 4" "3 2"; do
  set -- $shape
  export SA_NTT_ELOG=$1 SA_NTT_C=$2
  echo "== ELOG=$1 C=$2" >> gpurun_out/tune.log
  timeout 600 python -m pytest tests/test_gpu.py -q -x -k "test_ntt_matches_oracle and (18 or 19 or 20)" >> gpurun_out/tune.log 2>&1
  timeout 600 python bench.py --steps 50 --warmup 3 2>> gpurun_out/tune.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'elog': $1, 'c': $2, 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'single_ntt_us': d['single_ntt_us'], 'e2e': d['e2e']['value'], 'fri_ms': d['fri_commit_ms_2_20'], 'clocks': d['clocks']}))
" >> gpurun_out/tune.jsonl
done
cat gpurun_out/tune.jsonl
tail -20 gpurun_out/tune.log
