#!/bin/bash
# pipeline settings of sa_ntt_host: "streams chunk_MiB ramp" triples
mkdir -p gpurun_out; : > gpurun_out/e2e_sweep.txt
for cfg in "$@"; do
  set -- $cfg
  SA_HOST_STREAMS=$1 SA_HOST_CHUNK_MIB=$2 SA_HOST_RAMP=$3 timeout 300 python tools/e2e_sweep.py 2>&1 | tail -1 >> gpurun_out/e2e_sweep.txt
done
cat gpurun_out/e2e_sweep.txt
