#!/usr/bin/env python3
"""sa_ntt_host (pinned host buffers in and out) on 16 x 2^20: ms per call for the pipeline settings given
in the environment (SA_HOST_STREAMS, SA_HOST_CHUNK_MIB, SA_HOST_RAMP); checks the output once."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), os.path.join(ROOT, "oracle"), ROOT]
import numpy as np
import torch
import oracle as O
import sa_engine
eng = sa_engine.get_engine()  # (SA_B200_LIB selects an experiment build)
lib = eng.lib
LOG_N, BATCH = 20, 16
N = 1 << LOG_N
st = torch.cuda.current_stream()
hx = torch.randint(0, 1 << 62, (BATCH * N, 2), dtype=torch.int64).pin_memory()
hx[:, 1] &= (1 << 61) - 1
hy = torch.empty_like(hx).pin_memory()
root = sa_engine._limbs(O.primitive_nth_root(N))
call = lambda: lib.sa_ntt_host(hy.data_ptr(), hx.data_ptr(), LOG_N, root, 0, BATCH, ctypes.c_void_p(st.cuda_stream))
assert call() == 0
want = O.ntt_batch_np(O.primitive_nth_root(N), hx[-N:].numpy().view(np.uint64).reshape(1, N, 2)).reshape(-1, 2)
assert os.environ.get("SA_HOST_SKIP_NTT") or (hy[-N:].numpy().view(np.uint64) == want).all()
ts = []
for _ in range(8):
    t0 = time.perf_counter()
    assert call() == 0
    ts.append((time.perf_counter() - t0) * 1e3)
ms = sorted(ts)[len(ts) // 2]
print(json.dumps({"streams": os.environ.get("SA_HOST_STREAMS"), "chunk_mib": os.environ.get("SA_HOST_CHUNK_MIB"),
                  "ramp": os.environ.get("SA_HOST_RAMP"), "ms": round(ms, 3), "min_ms": round(min(ts), 3),
                  "GBps_each_way": round(BATCH * N * 16 / ms / 1e6, 1),
                  "butterflies_per_s": BATCH * (N // 2) * LOG_N / (ms * 1e-3)}))
