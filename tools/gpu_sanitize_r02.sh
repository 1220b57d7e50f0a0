#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_r02.py > gpurun_out/r02s_sanitize_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02s_sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_r02.py > gpurun_out/r02s_sanitize_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02s_sanitize_racecheck.log
tail -6 gpurun_out/r02s_sanitize_memcheck.log; tail -6 gpurun_out/r02s_sanitize_racecheck.log
