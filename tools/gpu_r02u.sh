#!/bin/bash
# start-up skew of the second CTA per SM (SA_NTT_SKEW_NS): do the two co-resident CTAs run in lock-step?
mkdir -p gpurun_out
: > gpurun_out/r02u_skew.jsonl
for cfg in "0 3" "3000 3" "6000 3" "10000 3" "13000 3" "13000 1" "18000 3" "0 3"; do
  set -- $cfg
  SA_BENCH_QUICK=1 SA_NTT_SKEW_NS=$1 SA_NTT_SKEW_PASSES=$2 timeout 300 python bench.py --steps 300 --warmup 3 >> gpurun_out/r02u_skew.jsonl 2>> gpurun_out/r02u.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02u_skew.jsonl'):
    d = json.loads(l); print(d['env'], round(d['ms_per_step'], 4), round(d['single_ntt_us'], 2))
PY
tail -3 gpurun_out/r02u.err
