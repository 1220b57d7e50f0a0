"""world_size-2 gloo test of the batch-sharding logic on CPU: every rank
transforms its slice through the engine interface (oracle test double here) and one all-gather
assembles the batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, log_n, q):
    sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    import oracle as O
    import sa_dist
    import sa_engine
    from fake_engine import OracleEngine
    sa_engine.set_engine(OracleEngine())
    n = 1 << log_n
    rng = np.random.default_rng(7)
    x = np.stack([rng.integers(0, 1 << 64, size=batch * n, dtype=np.uint64),
                  rng.integers(0, 0xCB80000000000000, size=batch * n, dtype=np.uint64)], axis=1)
    w = O.primitive_nth_root(n)
    full = sa_dist.sharded_ntt(x, log_n, w)
    local = sa_dist.sharded_ntt(x, log_n, w, gather=False)
    want = np.concatenate([O.ntt_np(w, x[b * n:(b + 1) * n]) for b in range(batch)], axis=0)
    lo, hi = sa_dist.shard_range(batch, rank, world)
    ok = bool((full == want).all()) and bool((local == want[lo * n:hi * n]).all())
    back = sa_dist.sharded_ntt(full, log_n, w, inverse=True)
    ok = ok and bool((back == x).all())
    # the peer-buffer modes refuse to run without their buffers (and a mode's buffers must be of its kind)
    if batch % world == 0:
        for mode in ("p2p-store", "p2p-push", "p2p-copy", "nvls-store", "nvls-push"):
            try:
                sa_dist.sharded_ntt(x, log_n, w, assemble=mode)
                ok = False
            except ValueError as exc:
                ok = ok and "needs peers" in str(exc)
        try:
            sa_dist.sharded_ntt(x, log_n, w, assemble="nvls-store", peers=object())
            ok = False
        except (ValueError, AttributeError):
            pass
    roots = sa_dist.sharded_merkle_roots(x, n)
    ok = ok and roots == [O.merkle_root_np(x[b * n:(b + 1) * n]) for b in range(batch)]
    # independent FRI instances: every rank commits its shard, the transcripts are gathered
    import pickle
    import fri as F
    from hostmirror_loader import load_host_types
    T = load_host_types()
    m = 64
    f = F.Fri(T.field.generator(), T.field.primitive_nth_root(m), m, 2, 4)
    cws = [[T.fe(int(v[0]) | (int(v[1]) << 64)) for v in x[b * m:(b + 1) * m]] for b in range(batch)]
    got = sa_dist.sharded_fri_commit(cws, f)
    ok = ok and len(got) == batch
    for b in range(batch):
        ps = F.ProofStream()
        F.Fri(T.field.generator(), T.field.primitive_nth_root(m), m, 2, 4).commit(cws[b], ps)
        ok = ok and pickle.dumps(got[b]) == pickle.dumps(ps.objects)
        oroots, _, _ = O.fri_commit_np(x[b * m:(b + 1) * m], O.GENERATOR, O.primitive_nth_root(m), 2, 4)
        ok = ok and [o for o in got[b] if isinstance(o, bytes)] == oroots
    q.put((rank, ok, sa_engine.get_engine().calls[0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_sharded_ntt_two_ranks(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in results), results
    sizes = sorted(c[3] for _, _, c in results)  # each rank transformed only its slice
    assert sum(sizes) == batch and sizes[0] >= batch // 2


def test_shard_range_covers_batch():
    sys.path.insert(0, os.path.join(ROOT, "stark-anatomy_b200"))
    import sa_dist
    for batch in (0, 1, 7, 8, 16, 17):
        for world in (1, 2, 3, 8):
            rs = [sa_dist.shard_range(batch, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == batch
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
