#!/bin/bash
# sweep SA_MK_SHAPE candidates (SA_TUNE build in build_variants/): args = shapes ("none" = defaults)
mkdir -p gpurun_out
rm -f gpurun_out/merkle_shapes.log
for shape in "$@"; do
  echo "== $shape" | tee -a gpurun_out/merkle_shapes.log
  SA_MK_SHAPE="$shape" SA_LIB=build_variants/libsa_tune.so timeout 300 python tools/merkle_sweep.py ${SWEEP_LOGS:-14 15 16 17 18 19 20} 2>&1 | tee -a gpurun_out/merkle_shapes.log
done
