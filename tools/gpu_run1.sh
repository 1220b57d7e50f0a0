#!/bin/bash
# first GPU session: parity tests, microbench, bench, ncu launch list + full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
if grep -q "test_field_selftest FAILED\|selftest" gpurun_out/pytest_gpu.log && grep -q "failed" gpurun_out/pytest_gpu.log; then
  echo "field selftest failed: rebuilding with portable field arithmetic" >> gpurun_out/pytest_gpu.log
  python -c "import __graft_entry__ as g; g.build_cuda(force=True, extra_flags=('-DSA_PORTABLE_FIELD',))" >> gpurun_out/build.log 2>&1
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_portable.log 2>&1
fi
timeout 300 python tools/microbench.py > gpurun_out/microbench.jsonl 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_tile_kernel -s 6 -c 2 -o gpurun_out/prof_ntt python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_merkle_chunk -c 2 -o gpurun_out/prof_merkle python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
