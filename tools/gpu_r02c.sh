#!/bin/bash
# round 2, third GPU pass (1 GPU): new parity tests (subproduct tree, sa_ntt_multi, table cache, host-entry knobs),
# a4-a6 timings, Montgomery-reduction variants x PDL x tile shapes on the bench workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest_gpu.log
timeout 600 python tools/poly_sweep.py > gpurun_out/r02c_poly_sweep.jsonl 2> gpurun_out/r02c_poly_sweep.err
rm -f gpurun_out/r02c_ntt_variants.jsonl
run() {  # lib-variant, PDL, ELOG, C
  SA_BENCH_QUICK=1 SA_B200_LIB=$PWD/build_variants/libsa_v$1.so SA_NTT_PDL=$2 SA_NTT_ELOG=$3 SA_NTT_C=$4 timeout 300 python bench.py --steps 300 --warmup 3 2>> gpurun_out/r02c_variants.err >> gpurun_out/r02c_ntt_variants.jsonl
}
run 1 0 0 0; run 1 1 0 0; run 2 0 0 0; run 2 1 0 0
run 2 1 4 2; run 2 1 3 4; run 2 1 3 2; run 2 1 4 8; run 1 1 4 2
timeout 900 python bench.py --steps 300 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -4 gpurun_out/r02c_pytest_gpu.log; cat gpurun_out/r02c_poly_sweep.jsonl; tail -3 gpurun_out/r02c_poly_sweep.err; cat gpurun_out/r02c_ntt_variants.jsonl | cut -c1-400; cut -c1-300 gpurun_out/r02c_bench.json
