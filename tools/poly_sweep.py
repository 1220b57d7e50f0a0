#!/usr/bin/env python3
"""Timings of sa_zerofier / sa_interpolate / sa_poly_eval on the GPU, device-resident inputs:
   python tools/poly_sweep.py [k ...]      -> one JSON line per k (ms per call, CUDA events, warm)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), os.path.join(ROOT, "oracle"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sa_engine  # noqa: E402

eng = sa_engine.get_engine()
ks = [int(a) for a in sys.argv[1:]] or [27, 284, 1024, 4096, 1 << 14, 1 << 15, 1 << 16, 1 << 18, 1 << 20]


def rand(seed, n):
    rng = np.random.default_rng(seed)
    return eng.upload(np.stack([rng.integers(0, 1 << 64, size=n, dtype=np.uint64),
                                rng.integers(0, 0xCB80000000000000, size=n, dtype=np.uint64)], axis=1).view(np.int64))


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for k in ks:
    dom, vals = rand(1, k), rand(2, k)
    reps = 20 if k <= 4096 else 3
    l0 = eng.launch_count()
    z = eng.zerofier(dom)
    launches_z = eng.launch_count() - l0
    line = {"k": k, "zerofier_ms": timed(lambda: eng.zerofier(dom), reps), "zerofier_launches": launches_z}
    if k <= 1 << 20:
        l0 = eng.launch_count()
        poly = eng.interpolate(dom, vals)
        line["interpolate_launches"] = eng.launch_count() - l0
        line["interpolate_ms"] = timed(lambda: eng.interpolate(dom, vals), reps)
        if k <= 1 << 17:
            line["poly_eval_horner_ms"] = timed(lambda: eng.poly_eval(poly, dom, mode=1), reps)
        l0 = eng.launch_count()
        eng.poly_eval(poly, dom, mode=2)
        line["poly_eval_walk_launches"] = eng.launch_count() - l0
        line["poly_eval_walk_ms"] = timed(lambda: eng.poly_eval(poly, dom, mode=2), reps)
        line["poly_eval_auto_ms"] = timed(lambda: eng.poly_eval(poly, dom), reps)
        ok = bool((eng.poly_eval(poly, dom, mode=2) == vals).all().item())
        line["interpolant_takes_values"] = ok
    print(json.dumps(line), flush=True)
