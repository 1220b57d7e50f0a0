#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_merkle -c 80 --csv --log-file gpurun_out/launches_merkle.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
