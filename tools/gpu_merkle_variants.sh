#!/bin/bash
# sweep the Merkle / FRI-round latency over experiment builds in build_variants/
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/merkle_variants.log
  SA_LIB=build_variants/libsa_$v.so timeout 300 python tools/merkle_sweep.py 2>&1 | tee -a gpurun_out/merkle_variants.log
done
