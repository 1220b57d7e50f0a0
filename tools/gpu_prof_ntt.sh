#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 100 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_tile_kernel -s 6 -c 2 -o gpurun_out/prof_ntt3 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['single_ntt_us'], d['fri_commit_ms_2_20'])"
