"""sa_dist -- multi-GPU use of the engine.

The hot path shards at batch level: independent transforms (one per polynomial / register /
column) and independent FRI instances have no data dependence, so every rank (one process per
GPU, ``torch.distributed``) runs its slice of the batch with no communication during compute.
What is left is ASSEMBLY: the caller wants the transformed batch on every rank.  Several ways to do
it, all bit-exact (``assemble=`` of ``sharded_ntt``):

  "p2p-store"  the product.  Every rank maps the other ranks' output buffers into its own device's address
               space (CUDA IPC over NVLink / NVSwitch, ``PeerBuffers``) and the LAST pass of its transforms stores each result tile
               to all of them (``sa_ntt_multi``): compute and assembly are one kernel, the NVLink
               writes overlap the butterflies tile by tile, no gather pass exists.
  "p2p-push"   transforms go to the local buffer; as soon as transform i is done ONE push kernel (``sa_push``:
               one read, a fully coalesced store per peer) sends it to all peers on a high-priority side
               stream while transform i+1 computes.
  "nvls-store" / "nvls-push"  the same two with ONE multicast address instead of 7 peer addresses (``McastBuffers``
               over an NVLS multicast object, ``sa_ntt_mcast`` / ``sa_push_mcast``): a store leaves the GPU once and
               the NVSwitch replicates it to every rank.
  "p2p-copy"   the same with the copy engines (one ``cudaMemcpyAsync`` and stream per peer).
  "nccl-pipelined"  cyclic ownership (rank r owns transforms r, r + world, ...), so chunk i of every rank
               is one in-place ``all_gather_into_tensor`` that runs on a side stream under chunk i+1.
  "nccl"       round 1's baseline: transform the contiguous slice, then ONE all-gather (also the only mode for
               ragged batches and for the gloo backend of the CPU tests).

A single transform or a single FRI commit is not split: FRI rounds are sequential through the host
Fiat-Shamir challenge.  Independent FRI instances shard like transforms (``sharded_fri_commit``).
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


def _rank_world(group=None):
    dist = _dist()
    if dist is None:
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


# --------------------------------------------------------------------------- peer-mapped buffers
class PeerBuffers:
    """``count`` symmetric device buffers of ``nelems`` field elements, each mapped on every rank.

    Rank r allocates its buffers through the C library (``sa_peer_alloc``: ``cudaMalloc`` + CUDA IPC handle),
    the 64-byte handles travel through ``all_gather_object`` and every other rank opens them with ITS device
    current (``sa_peer_open``: ``cudaIpcOpenMemHandle`` with lazy peer access), which maps the memory into that
    device's address space over NVLink / NVSwitch so that its kernels and copy engines can write it.
    ``local[k]`` is this rank's k-th buffer as a tensor, ``ptrs[k][q]`` rank q's k-th buffer as a raw device
    pointer valid in THIS process.  Two buffers alternate between calls so that a rank may still read call k's
    result while another rank already writes call k+1 (see ``sharded_ntt``).
    """

    def __init__(self, nelems, group=None, count=2):
        import ctypes
        import torch
        dist = _dist()
        eng = sa_engine.get_engine()
        lib = eng.lib
        self.group = group
        self.rank, self.world = _rank_world(group)
        self.nelems = nelems
        self.eng = eng
        self._own, self._opened = [], []
        self.local, self.ptrs = [], []
        eng._stream()  # (pins the current device)
        for k in range(count):
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            eng._check(lib.sa_peer_alloc(ctypes.byref(ptr), 16 * nelems, handle))
            self._own.append(ptr.value)
            self.local.append(eng.wrap_pointer(ptr.value, nelems))
            row = [None] * self.world
            row[self.rank] = ptr.value
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
                for q, h in enumerate(handles):
                    if q == self.rank:
                        continue
                    peer = ctypes.c_void_p()
                    eng._check(lib.sa_peer_open(ctypes.byref(peer), h))
                    self._opened.append(peer.value)
                    row[q] = peer.value
            self.ptrs.append(row)
        self.turn = 0
        self.flag = torch.zeros(1, dtype=torch.int32, device=eng.device)
        self.side = None
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=group)  # every rank has opened every handle before anybody writes

    def next(self):
        """(local buffer as a tensor, [rank q's buffer as a device pointer for q in range(world)]) of this call"""
        k = self.turn
        self.turn = (self.turn + 1) % len(self.local)
        return self.local[k], self.ptrs[k]

    def fence(self):
        """stream-ordered barrier across the ranks (a 4-byte NCCL all-reduce on the current stream): when it
        has passed on this rank, every rank has finished the kernels / copies it enqueued before its own"""
        if self.world > 1:
            _dist().all_reduce(self.flag, group=self.group)

    def streams(self):
        import torch
        if self.side is None:
            self.side = [torch.cuda.Stream() for _ in range(self.world)]
        return self.side

    def push_stream(self):
        """high-priority side stream of the push kernels (their CTAs take the SM slots the compute kernel frees)"""
        import torch
        if getattr(self, "_push", None) is None:
            self._push = torch.cuda.Stream(priority=-1)
        return self._push

    def close(self):
        """unmap the peers' buffers and free the own ones (every rank, after a barrier: nobody may still write)"""
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            _dist().barrier(group=self.group)
        lib = self.eng.lib
        for p in self._opened:
            lib.sa_peer_close(p)
        self._opened = []
        if self.world > 1:
            _dist().barrier(group=self.group)
        self.local = []
        for p in self._own:
            lib.sa_peer_free(p)
        self._own = []


class McastBuffers:
    """``count`` symmetric buffers of ``nelems`` field elements bound to an NVLS multicast object.

    ``torch.distributed._symmetric_memory`` does the plumbing (cuMemCreate + handle exchange + cuMulticastCreate /
    cuMulticastBindMem): ``local[k]`` is this rank's k-th buffer, ``mc[k]`` the multicast address of the k-th
    buffer set -- a store to ``mc[k] + off`` leaves this GPU ONCE and the NVSwitch writes it at ``off`` of every
    rank's buffer, this rank's included.  Raises if the box has no multicast support (no NVSwitch / driver without
    fabric support); callers fall back to ``PeerBuffers``.  Same ``next`` / ``fence`` / ``close`` protocol.
    """

    def __init__(self, nelems, group=None, count=2):
        import torch
        import torch.distributed._symmetric_memory as symm
        dist = _dist()
        if dist is None:
            raise RuntimeError("McastBuffers needs an initialised process group")
        eng = sa_engine.get_engine()
        self.eng, self.group, self.nelems = eng, group, nelems
        self.rank, self.world = _rank_world(group)
        grp = group if group is not None else dist.group.WORLD
        self.local, self.mc, self.ptrs, self._handles = [], [], [], []
        for _ in range(count):
            t = symm.empty((nelems, 2), dtype=torch.int64, device=eng.device)
            hdl = symm.rendezvous(t, grp)
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            if mc == 0:
                raise RuntimeError("no multicast (NVLS) support on this box: symmetric memory has no multicast_ptr")
            t.zero_()
            self.local.append(t)
            self.mc.append(mc)
            self.ptrs.append([int(x) for x in hdl.buffer_ptrs])
            self._handles.append(hdl)
        self.turn = 0
        self.flag = torch.zeros(1, dtype=torch.int32, device=eng.device)
        self._push = None
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def next(self):
        """(local buffer as a tensor, multicast address of the same buffer set) of this call"""
        k = self.turn
        self.turn = (self.turn + 1) % len(self.local)
        return self.local[k], self.mc[k]

    def fence(self):
        if self.world > 1:
            _dist().all_reduce(self.flag, group=self.group)

    def push_stream(self):
        import torch
        if self._push is None:
            self._push = torch.cuda.Stream(priority=-1)
        return self._push

    def close(self):
        import torch
        torch.cuda.synchronize()
        _dist().barrier(group=self.group)
        self.local, self.mc, self.ptrs, self._handles = [], [], [], []


# ------------------------------------------------------------------------------- sharded transforms
def owner_cyclic(b, world):
    return b % world


def sharded_ntt(vectors, log_n, root, inverse=False, gather=True, group=None, assemble="nccl", peers=None):
    """Transform a batch of B independent 2^log_n-point vectors, B split across the ranks
    (code/ntt.py:3-30 per transform).

    vectors: engine vector of B*n elements (every rank passes the same batch, or at least its own
    slice filled in).  Returns the full batch on every rank when ``gather`` (see the module docstring
    for ``assemble``), else only this rank's transformed slice.  The p2p modes need ``peers`` (a
    ``PeerBuffers`` of B*n elements, reusable across calls) and an even split; the returned tensor is
    one of its buffers and is overwritten by the call after the next one.
    """
    eng = sa_engine.get_engine()
    n = 1 << log_n
    batch = eng.length(vectors) // n
    rank, world = _rank_world(group)
    if not gather or world == 1 or assemble == "nccl" or batch % world != 0:
        lo, hi = shard_range(batch, rank, world)
        local = eng.ntt(eng.slice(vectors, lo * n, hi * n), log_n, root, inverse=inverse, batch=hi - lo) \
            if hi > lo else eng.empty(0)
        if not gather or world == 1:
            return local
        return all_gather_vectors(local, [shard_range(batch, r, world) for r in range(world)], n, group)
    per = batch // world
    if assemble == "nccl-pipelined":
        return _ntt_nccl_pipelined(eng, vectors, log_n, root, inverse, n, per, rank, world, group)
    if peers is None:
        raise ValueError("assemble=%r needs peers=PeerBuffers(batch * n) / McastBuffers(batch * n)" % assemble)
    lo = rank * per
    mine = eng.slice(vectors, lo * n, (lo + per) * n)
    local, bufs = peers.next()
    if assemble in ("nvls-store", "nvls-push"):
        if not isinstance(peers, McastBuffers):
            raise ValueError("assemble=%r needs peers=McastBuffers(batch * n)" % assemble)
        if assemble == "nvls-store":
            eng.ntt_mcast(bufs, local, lo * n, mine, log_n, root, inverse=inverse, batch=per)
        else:
            _ntt_mcast_push(eng, peers, local, bufs, mine, log_n, root, inverse, n, per, lo)
    elif not isinstance(peers, PeerBuffers):
        raise ValueError("assemble=%r needs peers=PeerBuffers(batch * n)" % assemble)
    elif assemble == "p2p-store":
        outs = [bufs[rank]] + [bufs[q] for q in range(world) if q != rank]
        eng.ntt_multi(outs, lo * n, mine, log_n, root, inverse=inverse, batch=per)
    elif assemble == "p2p-copy":
        _ntt_p2p_copy(eng, peers, local, bufs, mine, log_n, root, inverse, n, per, lo, rank, world)
    elif assemble == "p2p-push":
        _ntt_p2p_push(eng, peers, local, bufs, mine, log_n, root, inverse, n, per, lo, rank, world)
    else:
        raise ValueError("unknown assemble mode %r" % assemble)
    peers.fence()
    return local


def _ntt_p2p_copy(eng, peers, local, bufs, mine, log_n, root, inverse, n, per, lo, rank, world):
    import ctypes
    import torch
    side = peers.streams()
    main = torch.cuda.current_stream()
    for i in range(per):
        dst = local[(lo + i) * n:(lo + i + 1) * n]
        eng.ntt_into(dst, mine[i * n:(i + 1) * n], log_n, root, inverse=inverse)
        ev = torch.cuda.Event()
        ev.record(main)
        off = 16 * (lo + i) * n
        for q in range(world):
            if q == rank:
                continue
            side[q].wait_event(ev)  # the copy engine pushes transform i to rank q under transform i + 1
            eng._check(eng.lib.sa_copy_async(bufs[q] + off, bufs[rank] + off, 16 * n,
                                             ctypes.c_void_p(side[q].cuda_stream)))
    for q in range(world):
        if q != rank:
            main.wait_stream(side[q])


def _ntt_p2p_push(eng, peers, local, bufs, mine, log_n, root, inverse, n, per, lo, rank, world):
    import ctypes
    import torch
    side = peers.push_stream()
    main = torch.cuda.current_stream()
    others = [q for q in range(world) if q != rank]
    for i in range(per):
        dst = local[(lo + i) * n:(lo + i + 1) * n]
        eng.ntt_into(dst, mine[i * n:(i + 1) * n], log_n, root, inverse=inverse)
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        off = 16 * (lo + i) * n
        dsts = (ctypes.c_void_p * len(others))(*[bufs[q] + off for q in others])
        eng._check(eng.lib.sa_push(dsts, len(others), bufs[rank] + off, 16 * n, ctypes.c_void_p(side.cuda_stream)))
    main.wait_stream(side)


def _ntt_mcast_push(eng, peers, local, mc, mine, log_n, root, inverse, n, per, lo):
    import ctypes
    import torch
    side = peers.push_stream()
    main = torch.cuda.current_stream()
    for i in range(per):
        dst = local[(lo + i) * n:(lo + i + 1) * n]
        eng.ntt_into(dst, mine[i * n:(i + 1) * n], log_n, root, inverse=inverse)
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        off = 16 * (lo + i) * n  # one multimem store per 16 bytes: the switch delivers it to every rank
        eng._check(eng.lib.sa_push_mcast(ctypes.c_void_p(mc + off), ctypes.c_void_p(dst.data_ptr()), 16 * n,
                                         ctypes.c_void_p(side.cuda_stream)))
    main.wait_stream(side)


_COMM_STREAMS = {}


def _ntt_nccl_pipelined(eng, vectors, log_n, root, inverse, n, per, rank, world, group):
    import torch
    dist = _dist()
    out = eng.empty(per * world * n)
    main = torch.cuda.current_stream()
    comm = _COMM_STREAMS.get(main.device)
    if comm is None:
        comm = _COMM_STREAMS[main.device] = torch.cuda.Stream()
    works = []
    for i in range(per):
        b = i * world + rank  # cyclic ownership: chunk i of every rank is contiguous in batch order
        dst = out[b * n:(b + 1) * n]
        eng.ntt_into(dst, vectors[b * n:(b + 1) * n], log_n, root, inverse=inverse)
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            works.append(dist.all_gather_into_tensor(out[i * world * n:(i + 1) * world * n], dst, group=group,
                                                     async_op=True))
    for wk in works:
        wk.wait()
    main.wait_stream(comm)
    return out


def all_gather_vectors(local, ranges, n, group=None):
    """one all-gather of per-rank slices (padded to the largest slice) -> concatenated batch"""
    import torch
    dist = _dist()
    world = len(ranges)
    longest = max(hi - lo for lo, hi in ranges) * n
    is_torch = isinstance(local, torch.Tensor)
    t = local if is_torch else torch.from_numpy(np.ascontiguousarray(local).view(np.int64))
    even = all(hi - lo == ranges[0][1] - ranges[0][0] for lo, hi in ranges)
    if even and t.shape[0] == longest:
        pad = t.contiguous()
    else:
        pad = torch.zeros((longest, 2), dtype=torch.int64, device=t.device)
        pad[:t.shape[0]] = t
    out = torch.empty((world * longest, 2), dtype=torch.int64, device=t.device)
    if t.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.cat(parts, dim=0)
    if even:
        full = out
    else:
        full = torch.cat([out[r * longest:r * longest + (hi - lo) * n] for r, (lo, hi) in enumerate(ranges)], dim=0)
    return full if is_torch else full.numpy().view(np.uint64)


# ---------------------------------------------------------------------------- Merkle / FRI instances
def sharded_merkle_roots(vectors, n, group=None):
    """Merkle roots (code/merkle.py:13-14) of a batch of B independent n-element codewords, B split
    across the ranks; the 64-byte roots are all-gathered (B * 64 bytes in total).  Returns list[bytes]."""
    import torch
    eng = sa_engine.get_engine()
    dist = _dist()
    batch = eng.length(vectors) // n
    rank, world = _rank_world(group)
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


def sharded_fri_commit(codewords, fri, make_stream=None, group=None):
    """``Fri.commit`` (code/fri.py:56-96) of B INDEPENDENT instances, B split across the ranks.

    codewords : list of B codewords of ``fri.domain_length`` elements each (lists of FieldElement,
                DeviceCodeword objects or engine vectors wrapped by the caller); every rank passes the same list
                and only reads its own shard.
    fri       : the drop-in ``fri.Fri`` instance (same parameters for every instance).
    make_stream(b) -> the proof stream object of instance b (default: a fresh ``ProofStream``); every
                instance has its own transcript, so its challenges depend on its own roots only.
    Returns, on every rank, ``[objects_0, ..., objects_{B-1}]``: what instance b's commit pushed into its
    proof stream (the round roots, then the last codeword) -- gathered with one ``all_gather_object``
    (a few KB per instance; the big layers and trees stay on the rank that owns the instance).
    Rounds of ONE commit stay sequential (host Fiat-Shamir); nothing is split inside an instance.
    """
    import sa_host
    dist = _dist()
    rank, world = _rank_world(group)
    batch = len(codewords)
    lo, hi = shard_range(batch, rank, world)
    mine = []
    for b in range(lo, hi):
        ps = make_stream(b) if make_stream is not None else sa_host.ip.ProofStream()
        before = len(ps.objects)
        fri.commit(codewords[b], ps)
        mine.append(ps.objects[before:])
    if world == 1:
        return mine
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    return [objs for part in parts for objs in part]
