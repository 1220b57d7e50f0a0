#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/size_sweep.py > gpurun_out/size_sweep.jsonl 2> gpurun_out/size_sweep.err
timeout 600 python bench.py --steps 300 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/size_sweep.jsonl; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['int_roofline'])"
