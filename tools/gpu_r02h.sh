#!/bin/bash
# 1-GPU experiments: single-column CTAs for lone transforms, Merkle kernel occupancy variants
mkdir -p gpurun_out
rm -f gpurun_out/r02h_small_tiles.jsonl gpurun_out/r02h_merkle_minb.txt
for st in 0 300 600 1200; do
  SA_BENCH_QUICK=1 SA_NTT_SMALL_TILES=$st timeout 300 python bench.py --steps 300 --warmup 3 2>> gpurun_out/r02h.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'small_tiles_max': $st, 'ms_per_step': d['ms_per_step'], 'single_ntt_us': d['single_ntt_us']}))" >> gpurun_out/r02h_small_tiles.jsonl
done
SA_NTT_SMALL_TILES=600 timeout 600 python -m pytest tests/test_gpu.py -q -x -k "ntt_matches_oracle or three_pass or golden_digest or ntt_multi" > gpurun_out/r02h_pytest_small_tiles.log 2>&1
for lib in default build_variants/libsa_mk3.so build_variants/libsa_mk4.so; do
  echo "== $lib" >> gpurun_out/r02h_merkle_minb.txt
  if [ "$lib" = default ]; then timeout 300 python tools/merkle_sweep.py 14 16 17 18 19 20 >> gpurun_out/r02h_merkle_minb.txt 2>> gpurun_out/r02h.err
  else SA_LIB=$lib timeout 300 python tools/merkle_sweep.py 14 16 17 18 19 20 >> gpurun_out/r02h_merkle_minb.txt 2>> gpurun_out/r02h.err; fi
done
cat gpurun_out/r02h_small_tiles.jsonl; tail -2 gpurun_out/r02h_pytest_small_tiles.log; cat gpurun_out/r02h_merkle_minb.txt | cut -c1-260; tail -3 gpurun_out/r02h.err
