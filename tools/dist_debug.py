#!/usr/bin/env python3
"""step-by-step multi-GPU diagnosis (run under torchrun, CUDA_LAUNCH_BLOCKING=1): prints what it is about to do"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), os.path.join(ROOT, "oracle"), ROOT]
import numpy as np, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import oracle as O, sa_dist, sa_engine
eng = sa_engine.get_engine()
rank, world = dist.get_rank(), dist.get_world_size()
stage = sys.argv[1] if len(sys.argv) > 1 else "all"
def say(msg):
    print("[rank %d] %s" % (rank, msg), flush=True)
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
batch = 4 * world if log_n < 20 else 16
n = 1 << log_n
rng = np.random.default_rng(3)
x = np.stack([rng.integers(0, 1 << 64, size=batch * n, dtype=np.uint64), rng.integers(0, 0xCB80000000000000, size=batch * n, dtype=np.uint64)], axis=1)
w = O.primitive_nth_root(n)
vx = eng.upload(x.view(np.int64))
want = O.ntt_batch_np(w, x.reshape(batch, n, 2)).reshape(-1, 2)
torch.cuda.synchronize(); say("inputs ready")
def check(name, full):
    torch.cuda.synchronize()
    ok = bool((full.cpu().numpy().view(np.uint64) == want).all())
    say("%s -> %s" % (name, "ok" if ok else "MISMATCH"))
if stage in ("all", "nccl"):
    say("nccl mode"); check("nccl", sa_dist.sharded_ntt(vx, log_n, w, assemble="nccl"))
    say("nccl-pipelined mode"); check("nccl-pipelined", sa_dist.sharded_ntt(vx, log_n, w, assemble="nccl-pipelined"))
if stage in ("all", "peers"):
    say("PeerBuffers"); peers = sa_dist.PeerBuffers(batch * n); torch.cuda.synchronize(); say("PeerBuffers ok")
    say("copy engine into a peer buffer"); q = (rank + 1) % world
    import ctypes
    eng._check(eng.lib.sa_copy_async(peers.ptrs[0][q] + 128 * rank, vx.data_ptr(), 128, None)); torch.cuda.synchronize(); dist.barrier(); say("peer copy ok")
    say("nccl after PeerBuffers"); check("nccl", sa_dist.sharded_ntt(vx, log_n, w, assemble="nccl"))
    say("p2p-copy"); check("p2p-copy", sa_dist.sharded_ntt(vx, log_n, w, assemble="p2p-copy", peers=peers))
    say("p2p-store"); check("p2p-store", sa_dist.sharded_ntt(vx, log_n, w, assemble="p2p-store", peers=peers))
    say("p2p-store again"); check("p2p-store", sa_dist.sharded_ntt(vx, log_n, w, assemble="p2p-store", peers=peers))
say("done")
dist.destroy_process_group()
