#!/bin/bash
# Merkle kernel register bound (2 CTAs/SM restored) and the 3-CTAs/SM variant
mkdir -p gpurun_out
rm -f gpurun_out/r02i_merkle_minb.txt
for lib in default build_variants/libsa_mk3.so; do
  echo "== $lib" >> gpurun_out/r02i_merkle_minb.txt
  if [ "$lib" = default ]; then timeout 300 python tools/merkle_sweep.py 12 14 16 17 18 19 20 >> gpurun_out/r02i_merkle_minb.txt 2>> gpurun_out/r02i.err
  else SA_LIB=$lib timeout 300 python tools/merkle_sweep.py 12 14 16 17 18 19 20 >> gpurun_out/r02i_merkle_minb.txt 2>> gpurun_out/r02i.err; fi
done
timeout 600 python -m pytest tests/test_gpu.py -q -x -k "merkle or fri or peer_buffers" > gpurun_out/r02i_pytest.log 2>&1
cat gpurun_out/r02i_merkle_minb.txt | cut -c1-200; tail -2 gpurun_out/r02i_pytest.log; tail -3 gpurun_out/r02i.err
