#!/usr/bin/env python3
"""torchrun check of sa_dist on real GPUs: every assembly mode of sharded_ntt (fused peer stores, copy-engine
pushes, pipelined and plain NCCL all-gather) against the oracle, with device timing (max over ranks), and
sharded_fri_commit (independent FRI instances).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [log_n] [batch]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), os.path.join(ROOT, "oracle"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import oracle as O  # noqa: E402
import sa_dist  # noqa: E402
import sa_engine  # noqa: E402

eng = sa_engine.get_engine()
rank, world = dist.get_rank(), dist.get_world_size()
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = 1 << log_n
rng = np.random.default_rng(3)
x = np.stack([rng.integers(0, 1 << 64, size=batch * n, dtype=np.uint64),
              rng.integers(0, 0xCB80000000000000, size=batch * n, dtype=np.uint64)], axis=1)
w = O.primitive_nth_root(n)
vx = eng.upload(x.view(np.int64))
O.lib().so_set_threads(max(1, (os.cpu_count() or 8) // world))
check = sorted({0, batch // 2, batch - 1, (rank * 5 + 3) % batch})
want = {b: O.ntt_np(w, x[b * n:(b + 1) * n], parallel=True) for b in check}
report = {"rank": rank, "world": world, "log_n": log_n, "batch": batch, "modes": {}}
try:
    peers = sa_dist.PeerBuffers(batch * n)
except Exception as exc:
    peers = None
    report["peer_buffers_error"] = repr(exc)[:400]
MODES = os.environ.get("SA_DIST_MODES", "nccl,nccl-pipelined,p2p-copy,p2p-store,p2p-push,nvls-store,nvls-push").split(",")
mcast = None
if any(m.startswith("nvls") for m in MODES):
    try:
        mcast = sa_dist.McastBuffers(batch * n)
    except Exception as exc:
        report["mcast_buffers_error"] = repr(exc)[:400]
all_peers = peers
for mode in MODES:  # (the plain ones first: a fault in a peer mode cannot hide them)
    peers = mcast if mode.startswith("nvls") else all_peers
    if mode.startswith(("p2p", "nvls")) and peers is None:
        continue
    try:
        for _ in range(3):
            full = sa_dist.sharded_ntt(vx, log_n, w, assemble=mode, peers=peers)
        torch.cuda.synchronize()
        dist.barrier()
        got = full.cpu().numpy().view(np.uint64)
        ok = all((got[b * n:(b + 1) * n] == want[b]).all() for b in check)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        torch.cuda.synchronize()
        dist.barrier()
        ev0.record()
        for _ in range(reps):
            full = sa_dist.sharded_ntt(vx, log_n, w, assemble=mode, peers=peers)
        ev1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1) / reps], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the inverse through the same mode must give the input back on every rank
        back = sa_dist.sharded_ntt(full.clone(), log_n, w, inverse=True, assemble=mode, peers=peers)
        torch.cuda.synchronize()
        ok = ok and bool((back.cpu().numpy().view(np.uint64)[check[-1] * n:(check[-1] + 1) * n] == x[check[-1] * n:(check[-1] + 1) * n]).all())
        report["modes"][mode] = {"ok": bool(ok), "ms_per_call_max_over_ranks": float(t.item())}
    except Exception as exc:
        report["modes"][mode] = {"error": repr(exc)[:400]}
        torch.cuda.synchronize()

# independent FRI instances, one transcript each (2^16 codewords, 4 per rank)
report["env"] = {k: v for k, v in os.environ.items() if k.startswith(("SA_NTT_PEER", "SA_PUSH"))}
try:
    if os.environ.get("SA_DIST_SKIP_FRI") == "1":
        raise RuntimeError("skipped")
    import hashlib
    import pickle
    import sa_host
    import sa_marshal
    import fri as F
    from sa_devlist import DeviceCodeword
    field = sa_host.algebra.Field.main()
    FE = sa_host.algebra.FieldElement
    m, inst = 1 << 16, 4 * world
    f = F.Fri(field.generator(), FE(O.primitive_nth_root(m), field), m, 4, 64)
    cws = [DeviceCodeword(vx[b * m:(b + 1) * m].contiguous(), None, field) for b in range(inst)]
    t0 = time.perf_counter()
    got = sa_dist.sharded_fri_commit(cws, f)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    b = (rank * 3 + 1) % inst
    oroots, _, layers = O.fri_commit_np(x[b * m:(b + 1) * m], O.GENERATOR, O.primitive_nth_root(m), 4, 64)
    ok = [o for o in got[b] if isinstance(o, bytes)] == oroots and [v.value for v in got[b][-1]] == O.from_np(layers[-1])
    report["sharded_fri_commit"] = {"ok": bool(ok), "instances": inst, "n": m, "seconds": dt,
                                    "transcripts_sha256": hashlib.sha256(pickle.dumps(got)).hexdigest()[:16]}
except Exception as exc:
    report["sharded_fri_commit"] = {"error": repr(exc)[:400]}
print("DIST_CHECK " + json.dumps(report), flush=True)
dist.destroy_process_group()
