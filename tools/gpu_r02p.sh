#!/bin/bash
# multi-point evaluation by the walk down the subproduct tree: parity tests, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -x -q -k "poly_eval_tree_walk or zerofier_and_interpolate or elementwise" 2>&1 | tail -15 > gpurun_out/r02p_pytest.log
cat gpurun_out/r02p_pytest.log
timeout 600 python tools/poly_sweep.py > gpurun_out/r02p_poly_sweep.jsonl 2> gpurun_out/r02p_poly_sweep.err
cat gpurun_out/r02p_poly_sweep.jsonl; tail -5 gpurun_out/r02p_poly_sweep.err
