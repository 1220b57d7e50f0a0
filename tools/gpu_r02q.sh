#!/bin/bash
# lone 2^20 transform: seven-column tiles (147 tiles, one per SM) against four-column tiles (256 tiles on 296 slots)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -q -x -k "test_ntt_matches_oracle or golden_digest or random_campaign or in_place" 2>&1 | tail -2 > gpurun_out/r02q_pytest.log; cat gpurun_out/r02q_pytest.log
: > gpurun_out/r02q_lone_tiles.jsonl
for c in 4 7 4 7; do
  SA_BENCH_QUICK=1 SA_NTT_LONE_C=$c timeout 300 python bench.py --steps 200 --warmup 3 >> gpurun_out/r02q_lone_tiles.jsonl 2>> gpurun_out/r02q.err
done
cat gpurun_out/r02q_lone_tiles.jsonl; tail -3 gpurun_out/r02q.err
