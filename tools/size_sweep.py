#!/usr/bin/env python3
"""Device-resident NTT throughput by transform size (batch sized to ~256 MiB): one JSON line per size."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), os.path.join(ROOT, "oracle"), ROOT]
import torch
import oracle as O
import sa_engine
eng = sa_engine.get_engine()
lib, dev = eng.lib, eng.device
st = torch.cuda.current_stream()
total = 1 << 24  # elements per step (256 MiB)
x = torch.randint(0, 1 << 62, (total, 2), dtype=torch.int64, device=dev)
y = torch.empty_like(x)
for log_n in (4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24):
    n = 1 << log_n
    batch = total // n
    root = sa_engine._limbs(O.primitive_nth_root(n))
    call = lambda: lib.sa_ntt(y.data_ptr(), x.data_ptr(), log_n, root, 0, batch, ctypes.c_void_p(st.cuda_stream))
    for _ in range(3):
        assert call() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record(st)
    for _ in range(reps):
        call()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bf = batch * (n // 2) * log_n
    print(json.dumps({"log_n": log_n, "batch": batch, "ms": ms, "butterflies_per_s": bf / (ms * 1e-3),
                      "GB_per_s_alg": 32 * total / (ms * 1e-3) / 1e9}))
