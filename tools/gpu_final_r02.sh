#!/bin/bash
# end-of-round evidence (1 GPU): GPU tests, smoke, both bench arms, ncu launch list of the bench command,
# ncu --set full of the dominant kernel and of the Merkle kernel (raw CSV summaries -> profiles/)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_reference_arm.json 2> gpurun_out/r02z_bench_ref.err
timeout 900 python bench.py --steps 500 --warmup 3 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
SA_BENCH_QUICK=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02z_ncu_bench.log 2>&1
SA_BENCH_QUICK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_tile_kernel -s 6 -c 2 -o gpurun_out/r02z_prof_ntt python bench.py --steps 2 --warmup 3 > gpurun_out/r02z_ncu_full.log 2>&1
ncu -i gpurun_out/r02z_prof_ntt.ncu-rep --page raw --csv > gpurun_out/r02z_ncu_ntt_tile_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02z_ncu_ntt_tile_raw.csv > gpurun_out/r02z_ncu_ntt_tile_summary.txt 2>&1
cat > /tmp/mk.py <<'PY'
import sys
sys.path[:0] = ["stark-anatomy_b200", "oracle", "."]
import numpy as np, torch, sa_engine
eng = sa_engine.get_engine()
rng = np.random.default_rng(1)
N = 1 << 20
cw = eng.upload(np.stack([rng.integers(0, 1 << 64, size=N, dtype=np.uint64), rng.integers(0, 0xCB80000000000000, size=N, dtype=np.uint64)], axis=1).view(np.int64))
for _ in range(3):
    t = eng.merkle_tree(cw)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_merkle_chunk -s 4 -c 1 -o gpurun_out/r02z_prof_merkle python /tmp/mk.py > gpurun_out/r02z_ncu_full2.log 2>&1
ncu -i gpurun_out/r02z_prof_merkle.ncu-rep --page raw --csv > gpurun_out/r02z_ncu_merkle_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02z_ncu_merkle_raw.csv > gpurun_out/r02z_ncu_merkle_chunk_summary.txt 2>&1
rm -f gpurun_out/r02z_prof_ntt.ncu-rep gpurun_out/r02z_prof_merkle.ncu-rep
tail -3 gpurun_out/r02z_pytest_gpu.log; tail -2 gpurun_out/r02z_smoke.log; cut -c1-300 gpurun_out/r02z_bench.json; cut -c1-200 gpurun_out/r02z_bench_reference_arm.json; head -30 gpurun_out/r02z_ncu_ntt_tile_summary.txt
