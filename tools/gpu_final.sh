#!/bin/bash
# end-of-round evidence: GPU tests, smoke, both bench arms, ncu launch list of the bench command,
# ncu --set full of the dominant kernel and of the Merkle kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_tile_kernel -s 6 -c 2 -o gpurun_out/prof_ntt_final python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_merkle_chunk -c 3 -o gpurun_out/prof_merkle_final python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json | cut -c1-300
