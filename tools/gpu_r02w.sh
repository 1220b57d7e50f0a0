#!/bin/bash
# decimal leaf encoding v2 (base-1e8 limbs, FP64 quotient estimates, multiply-built ASCII): parity, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -q -x -k "merkle or fri or dropin or config5 or field" 2>&1 | tail -3 > gpurun_out/r02w_pytest.log; cat gpurun_out/r02w_pytest.log
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02w_bench.json')); print(d['ms_per_step'], d['single_ntt_us'], d['fri_commit_ms_2_20'], d['fri_roofline']['ms_with_constant_challenge'], d['fri_roofline']['frac'], d['list_api']['fri_commit_device_list_in_s'])"
timeout 300 python tools/merkle_sweep.py 16 18 20 2>/dev/null | tail -4
