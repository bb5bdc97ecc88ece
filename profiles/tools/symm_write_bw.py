#!/usr/bin/env python
"""Write bandwidth into a peer's SYMMETRIC-memory buffer (torch.distributed._symmetric_memory) from a plain SM kernel and
from the copy engine, one direction and both directions at once.  torchrun --nproc-per-node 2 symm_write_bw.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
peer = (rank + 1) % world
for nbytes in (6520832, 13041664, 45645824):
    n = nbytes // 4
    buf = symm_mem.empty((n,), dtype=torch.float32, device=dev)
    buf.zero_()
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    remote = hdl.get_buffer(peer, (n,), torch.float32)
    plain_remote = None
    local = torch.ones(n, dtype=torch.float32, device=dev)
    local2 = torch.empty(n, dtype=torch.float32, device=dev)

    def timed(fn, active):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        if active:
            for _ in range(3):
                fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if active:
            for _ in range(20):
                fn()
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        return e0.elapsed_time(e1) / 20 * 1e3

    for name, fn in (("sm kernel (mul out=remote)", lambda: torch.mul(local, 1.0, out=remote)),
                     ("copy engine (remote.copy_)", lambda: remote.copy_(local, non_blocking=True)),
                     ("sm kernel local", lambda: torch.mul(local, 1.0, out=local2))):
        both = timed(fn, True)
        one = timed(fn, rank == 0)
        if rank == 0:
            print("%9d B  %-28s both directions %7.1f us %6.1f GB/s | rank 0 only %7.1f us %6.1f GB/s" % (
                nbytes, name, both, nbytes / both / 1e3, one, nbytes / one / 1e3), flush=True)
    del remote, hdl, buf
dist.barrier()
dist.destroy_process_group()
