#!/bin/bash
# round 2, third GPU pass (1 GPU): new parity tests (subproduct tree, sa_ntt_multi, table cache, host-entry knobs),
# a4-a6 timings, and the two Montgomery-reduction variants on the bench workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest_gpu.log
timeout 600 python tools/poly_sweep.py > gpurun_out/r02c_poly_sweep.jsonl 2> gpurun_out/r02c_poly_sweep.err
for v in 1 2; do
  SA_B200_LIB=$PWD/build_variants/libsa_v$v.so timeout 600 python bench.py --steps 300 --warmup 3 2> gpurun_out/r02c_bench_v$v.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'variant': $v, 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'single_ntt_us': d['single_ntt_us'], 'int_roofline': d['int_roofline'], 'fri_commit_ms': d['fri_commit_ms_2_20'], 'fri_roofline': d.get('fri_roofline'), 'list_api': d.get('list_api')}))" >> gpurun_out/r02c_montmul_variants.jsonl
done
tail -4 gpurun_out/r02c_pytest_gpu.log; cat gpurun_out/r02c_poly_sweep.jsonl; tail -3 gpurun_out/r02c_poly_sweep.err; cat gpurun_out/r02c_montmul_variants.jsonl | cut -c1-600
