#!/bin/bash
# final-code sweeps for the record: NTT throughput by transform size (1-, 2- and 3-pass plans), a4-a6 timings by point count
mkdir -p gpurun_out
timeout 400 python tools/size_sweep.py > gpurun_out/r02z_size_sweep.jsonl 2> gpurun_out/r02z_sweeps.err
timeout 300 python tools/poly_sweep.py > gpurun_out/r02z_poly_sweep.jsonl 2>> gpurun_out/r02z_sweeps.err
cut -c1-220 gpurun_out/r02z_size_sweep.jsonl; cut -c1-330 gpurun_out/r02z_poly_sweep.jsonl; tail -3 gpurun_out/r02z_sweeps.err
