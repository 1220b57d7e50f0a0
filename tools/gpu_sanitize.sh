#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu.py -q -x -k "test_ntt_matches_oracle and (1- or 5- or 9- or 10- or 11- or 13-) or test_merkle_tree_and_open or test_fri_round_and_fold or test_zerofier or test_elementwise" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu.py -q -x -k "test_ntt_matches_oracle and (5- or 10- or 12-) or test_fri_round_and_fold or test_zerofier" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck.log
tail -8 gpurun_out/sanitize_memcheck.log; tail -8 gpurun_out/sanitize_racecheck.log
