#!/usr/bin/env python3
"""torchrun check of sa_dist.sharded_ntt on real GPUs: NCCL all-gather, result vs oracle, timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), os.path.join(ROOT, "oracle"), ROOT]
import numpy as np, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import oracle as O, sa_dist, sa_engine
eng = sa_engine.get_engine()
log_n, batch = 20, 16
n = 1 << log_n
rng = np.random.default_rng(3)
x = np.stack([rng.integers(0, 1 << 64, size=batch * n, dtype=np.uint64), rng.integers(0, 0xCB80000000000000, size=batch * n, dtype=np.uint64)], axis=1)
w = O.primitive_nth_root(n)
vx = eng.upload(x.view(np.int64))
full = sa_dist.sharded_ntt(vx, log_n, w)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    full = sa_dist.sharded_ntt(vx, log_n, w)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
got = full.cpu().numpy().view(np.uint64)
r = dist.get_rank()
ok = all((got[b * n:(b + 1) * n] == O.ntt_np(w, x[b * n:(b + 1) * n], parallel=True)).all() for b in (0, batch // 2, batch - 1))
print(f"rank {r}: sharded_ntt batch={batch} 2^{log_n} ok={ok} {dt*1e3:.3f} ms per call incl. all-gather of {batch*n*16/2**20:.0f} MiB")
dist.destroy_process_group()
