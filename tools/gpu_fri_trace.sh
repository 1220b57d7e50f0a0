#!/bin/bash
mkdir -p gpurun_out
SA_FRI_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_trace.json 2> gpurun_out/fri_trace.log
tail -30 gpurun_out/fri_trace.log
