#!/bin/bash
# lone 2^20 transform: mixed grid of four- and three-column tiles (2 CTAs per SM exactly) against 256 four-column tiles
mkdir -p gpurun_out
: > gpurun_out/r02v_lone_mix.jsonl
for m in 0 1 2 0 1 2; do
  SA_BENCH_QUICK=1 SA_NTT_LONE_MIX=$m timeout 300 python bench.py --steps 100 --warmup 3 >> gpurun_out/r02v_lone_mix.jsonl 2>> gpurun_out/r02v.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02v_lone_mix.jsonl'):
    d = json.loads(l); print(d['env'], round(d['ms_per_step'], 4), round(d['single_ntt_us'], 2))
PY
tail -3 gpurun_out/r02v.err
