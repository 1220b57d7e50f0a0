#!/bin/bash
# round 2, fourth GPU pass (1 GPU): persistent FRI tail (tests both modes), per-round timeline, bench, and what the
# profiling tools do to the environment / to host<->kernel communication
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest_gpu.log
cat > /tmp/fri_trace.py <<'PY'
import sys, os, time, hashlib, pickle
sys.path[:0] = ["stark-anatomy_b200", "oracle", "."]
import numpy as np, torch, sa_engine, oracle as O
eng = sa_engine.get_engine()
N = 1 << 20
rng = np.random.default_rng(1)
cw = eng.upload(np.stack([rng.integers(0, 1 << 64, size=N, dtype=np.uint64), rng.integers(0, 0xCB80000000000000, size=N, dtype=np.uint64)], axis=1).view(np.int64))
w = O.primitive_nth_root(N)
def run(const):
    objs = []
    def on_root(r, root, want):
        objs.append(root)
        if not want: return None
        return 12345678901234567890 if const else O.sample(hashlib.shake_256(pickle.dumps(objs)).digest(32))
    eng.fri_commit(cw, 12, O.GENERATOR, w, on_root)
    return objs
for const in (False, True):
    run(const); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): run(const)
    torch.cuda.synchronize()
    print("fri_commit 2^20 %s challenge: %.4f ms, tail mode %d" % ("constant" if const else "python", (time.perf_counter() - t0) / 10 * 1e3, eng.lib.sa_fri_tail_mode()), flush=True)
os.environ["SA_FRI_TRACE_NOW"] = "1"
PY
timeout 300 python /tmp/fri_trace.py > gpurun_out/r02d_fri_modes.txt 2>&1
SA_FRI_PERSISTENT=0 timeout 300 python /tmp/fri_trace.py >> gpurun_out/r02d_fri_modes.txt 2>&1
SA_FRI_TRACE=1 timeout 300 python /tmp/fri_trace.py 2>&1 | tail -30 > gpurun_out/r02d_fri_timeline_tail.log
SA_FRI_TRACE=1 SA_FRI_PERSISTENT=0 timeout 300 python /tmp/fri_trace.py 2>&1 | tail -30 > gpurun_out/r02d_fri_timeline_per_round.log
# tools: environment they inject, and whether the probe falls back under them
(timeout 120 ncu env 2>&1 | grep -i -E "inject|nv_|nsight|profiler|preload" ; echo "--- sanitizer"; timeout 120 compute-sanitizer env 2>&1 | grep -i -E "inject|nv_|sanit|preload") > gpurun_out/r02d_tool_env.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02d_ncu_fri_launches.csv python /tmp/fri_trace.py > gpurun_out/r02d_ncu_fri.log 2>&1
timeout 900 python bench.py --steps 300 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02d_bench_reference.json 2> gpurun_out/r02d_bench_reference.err
tail -4 gpurun_out/r02d_pytest_gpu.log; cat gpurun_out/r02d_fri_modes.txt; cat gpurun_out/r02d_fri_timeline_tail.log | tail -13; cat gpurun_out/r02d_tool_env.txt; tail -3 gpurun_out/r02d_ncu_fri.log; cut -c1-200 gpurun_out/r02d_bench.json; cut -c1-400 gpurun_out/r02d_bench_reference.json
