#!/usr/bin/env python3
"""Device time of one Merkle tree build (sa_merkle_tree) and one fused FRI round (sa_fri_round) by
width, launches back to back on one stream (CUDA events), plus sa_fri_commit at 2^20 with a trivial
host challenge: one JSON line per width.  Usage: tools/merkle_sweep.py [log widths ...]"""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), os.path.join(ROOT, "oracle"), ROOT]
import torch
import oracle as O
import sa_engine
if os.environ.get("SA_LIB"):  # experiment builds (build_variants/)
    sa_engine.load_library.__defaults__ = (os.path.join(ROOT, os.environ["SA_LIB"]),)
eng = sa_engine.get_engine()
lib, dev = eng.lib, eng.device
st = torch.cuda.current_stream()
sp = ctypes.c_void_p(st.cuda_stream)
# parity spot check of whichever build was loaded (all launch shapes: 2^12 and 2^18 leaves)
import numpy as np
for chk in (12, 18):
    xs = torch.randint(0, 1 << 62, (1 << chk, 2), dtype=torch.int64, device=dev)
    xs[:, 1] &= (1 << 61) - 1
    t = eng.merkle_tree(xs)
    want = O.merkle_tree_np(xs.cpu().numpy().view(np.uint64))
    assert (t.cpu().numpy()[1:] == want[1:]).all(), "merkle parity FAILED at 2^%d" % chk
logs = [int(a) for a in sys.argv[1:]] or [1, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20]
P = O.P
for log_n in logs:
    n = 1 << log_n
    x = torch.randint(0, 1 << 62, (2 * n, 2), dtype=torch.int64, device=dev)
    x[:, 1] &= (1 << 61) - 1  # canonical (< p)
    tree = torch.empty((2 * n, 64), dtype=torch.uint8, device=dev)
    nxt = torch.empty((n, 2), dtype=torch.int64, device=dev)
    alpha, off = sa_engine._limbs(12345678901234567890123), sa_engine._limbs(O.GENERATOR if hasattr(O, "GENERATOR") else 3)
    om = sa_engine._limbs(O.primitive_nth_root(2 * n))
    calls = {
        "merkle_tree_us": lambda: lib.sa_merkle_tree(tree.data_ptr(), x.data_ptr(), n, sp),
        "fri_round_us": lambda: lib.sa_fri_round(nxt.data_ptr(), tree.data_ptr(), x.data_ptr(), 2 * n, alpha, off, om, sp),
    }
    rec = {"log_width": log_n}
    for name, call in calls.items():
        for _ in range(3):
            assert call() == 0
        torch.cuda.synchronize()
        reps = 50 if log_n <= 16 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            call()
        e1.record(st)
        torch.cuda.synchronize()
        rec[name] = round(e0.elapsed_time(e1) / reps * 1e3, 2)
    print(json.dumps(rec), flush=True)

# whole commit, 2^20, 12 rounds, host callback = constant challenge
n = 1 << 20
cw = torch.randint(0, 1 << 62, (n, 2), dtype=torch.int64, device=dev)
cw[:, 1] &= (1 << 61) - 1
om = O.primitive_nth_root(n)
ts = []
for it in range(12):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.fri_commit(cw, 12, 3, om, lambda r, root, want: 987654321987654321 + r)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"fri_commit_2_20_ms_min": round(min(ts[2:]), 4), "median": round(sorted(ts[2:])[len(ts[2:]) // 2], 4)}))
