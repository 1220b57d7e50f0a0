#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -q -x -k "test_ntt_matches_oracle and (19 or 20)" > gpurun_out/quick.log 2>&1; tail -1 gpurun_out/quick.log
timeout 600 python bench.py --steps 300 --warmup 3 2>> gpurun_out/quick.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['single_ntt_us'], d['roofline']['frac'])"
