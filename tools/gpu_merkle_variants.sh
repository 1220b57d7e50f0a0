#!/bin/bash
# sweep the Merkle / FRI-round latency over experiment builds in build_variants/ (args: build names)
mkdir -p gpurun_out
rm -f gpurun_out/merkle_variants.log
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/merkle_variants.log
  SA_LIB=build_variants/libsa_$v.so timeout 300 python tools/merkle_sweep.py ${SWEEP_LOGS:-8 12 14 16 17 18 19 20} 2>&1 | tee -a gpurun_out/merkle_variants.log
done
