#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_tile_kernel -s 6 -c 2 -o gpurun_out/prof_ntt2 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_merkle_chunk -c 3 -o gpurun_out/prof_merkle2 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
