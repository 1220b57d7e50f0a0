#!/bin/bash
# round 2, first GPU pass: parity suite on the reworked host side (device lists, LRU table cache, host-entry
# ramp fix), config 5 on the real engine, instruction-rate microbenchmarks, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r02a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest_gpu.log
timeout 1200 python tools/config5.py --out gpurun_out/r02a_config5.json > gpurun_out/r02a_config5.log 2>&1
timeout 300 ./tools/pipebench > gpurun_out/r02a_pipebench.jsonl 2>&1
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_pytest_gpu.log; head -c 1500 gpurun_out/r02a_config5.json; cat gpurun_out/r02a_pipebench.jsonl; cut -c1-400 gpurun_out/r02a_bench.json
