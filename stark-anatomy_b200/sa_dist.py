"""sa_dist -- multi-GPU use of the engine.

The hot path shards at batch level: independent transforms (one per polynomial / register /
column) and independent FRI instances have no data dependence, so every rank (one process per
GPU, ``torch.distributed``) runs its slice of the batch with no communication, and ONE
all-gather assembles the outputs where the caller wants them on every rank (NCCL over
NVLink / NVSwitch on GPUs; gloo for the CPU tests).  A single transform or a single FRI commit
is not split: FRI rounds are sequential through the host Fiat-Shamir challenge.
"""
import numpy as np

import sa_engine


def shard_range(batch, rank, world):
    """contiguous slice [lo, hi) of `batch` items owned by `rank` (sizes differ by at most one)"""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def sharded_ntt(vectors, log_n, root, inverse=False, gather=True, group=None):
    """Transform a batch of B independent 2^log_n-point vectors, B split across the ranks.

    vectors: engine vector of B*n elements (every rank passes the same batch, or at least its own
    slice filled in).  Returns the full batch on every rank when ``gather`` (one all-gather),
    else only this rank's transformed slice.  code/ntt.py:3-30 per transform.
    """
    eng = sa_engine.get_engine()
    dist = _dist()
    n = 1 << log_n
    batch = eng.length(vectors) // n
    rank = dist.get_rank(group) if dist else 0
    world = dist.get_world_size(group) if dist else 1
    lo, hi = shard_range(batch, rank, world)
    local = eng.ntt(eng.slice(vectors, lo * n, hi * n), log_n, root, inverse=inverse, batch=hi - lo) \
        if hi > lo else eng.empty(0)
    if not gather or world == 1:
        return local
    return all_gather_vectors(local, [shard_range(batch, r, world) for r in range(world)], n, group)


def all_gather_vectors(local, ranges, n, group=None):
    """one all-gather of per-rank slices (padded to the largest slice) -> concatenated batch"""
    import torch
    dist = _dist()
    eng = sa_engine.get_engine()
    world = len(ranges)
    longest = max(hi - lo for lo, hi in ranges) * n
    is_torch = isinstance(local, torch.Tensor)
    t = local if is_torch else torch.from_numpy(np.ascontiguousarray(local).view(np.int64))
    pad = torch.zeros((longest, 2), dtype=torch.int64, device=t.device)
    pad[:t.shape[0]] = t
    out = torch.empty((world * longest, 2), dtype=torch.int64, device=t.device)
    if t.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.cat(parts, dim=0)
    pieces = [out[r * longest:r * longest + (hi - lo) * n] for r, (lo, hi) in enumerate(ranges)]
    full = torch.cat(pieces, dim=0)
    return full if is_torch else full.numpy().view(np.uint64)


def sharded_merkle_roots(vectors, n, group=None):
    """Merkle roots (code/merkle.py:13-14) of a batch of B independent n-element codewords, B split
    across the ranks; the 64-byte roots are all-gathered (B * 64 bytes in total).  Returns list[bytes]."""
    import torch
    eng = sa_engine.get_engine()
    dist = _dist()
    batch = eng.length(vectors) // n
    rank = dist.get_rank(group) if dist else 0
    world = dist.get_world_size(group) if dist else 1
    lo, hi = shard_range(batch, rank, world)
    mine = [eng.tree_root(eng.merkle_tree(eng.slice(vectors, b * n, (b + 1) * n))) for b in range(lo, hi)]
    if world == 1:
        return mine
    longest = max(h - l for l, h in (shard_range(batch, r, world) for r in range(world)))
    buf = torch.zeros((longest, 64), dtype=torch.uint8)
    for i, r in enumerate(mine):
        buf[i] = torch.frombuffer(bytearray(r), dtype=torch.uint8)
    on_gpu = isinstance(vectors, torch.Tensor) and vectors.is_cuda
    if on_gpu:
        buf = buf.to(vectors.device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    roots = []
    for r in range(world):
        l, h = shard_range(batch, r, world)
        roots += [bytes(parts[r][i].cpu().numpy().tobytes()) for i in range(h - l)]
    return roots
