#!/bin/bash
# device-resident NTT step time (bench.py) for experiment builds in build_variants/ (args: build names)
mkdir -p gpurun_out
rm -f gpurun_out/bench_variants.log
for v in "$@"; do
  SA_B200_LIB=$PWD/build_variants/libsa_$v.so timeout 600 python bench.py --steps 300 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['value'], d.get('single_ntt_us'), d.get('fri_commit_ms_2_20'))" | tee -a gpurun_out/bench_variants.log
done
