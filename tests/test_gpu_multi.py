"""Two B200s, NCCL: env-id sharding + the observation all-gather, fused into the step kernel (peer stores over
NVLink into symmetric-memory buffers) and as a separate in-place ncclAllGather.  Every rank must end each step holding
exactly the rows a single GPU computes for the whole batch.  Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, E, steps, q):
    import torch.distributed as dist
    from crowdnav_b200.config import baseline_config
    from crowdnav_b200.sharded import ShardedVecEnv
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    from parity_util import random_actions
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg = baseline_config(3, n_envs=E, auto_reset=True)             # mixed behaviours, keyed by global env id
        try:
            senv = ShardedVecEnv(cfg, lambda c, o: CrowdNavVecEnv(c, device=rank, obs_out=o), dev, gather=mode)
        except RuntimeError as exc:
            if mode == "fused_mc" and "multicast" in str(exc):
                q.put((rank, "no-multicast"))
                return
            raise
        full = CrowdNavVecEnv(cfg, device=rank)                         # the whole batch on this GPU: the expectation
        rng = np.random.default_rng(7)
        senv.reset()
        full.reset()
        bad = 0
        for t in range(steps):
            a = torch.from_numpy(random_actions(rng, E)).to(dev)
            obs_all, r, d = senv.step(a[senv.lo:senv.hi].contiguous())
            senv.wait_gathered()
            fo, fr, fd = full.step(a)
            torch.cuda.synchronize()
            dist.barrier()
            if not (torch.equal(obs_all.view(torch.int32), fo.view(torch.int32)) and torch.equal(r, fr[senv.lo:senv.hi])
                    and torch.equal(d, fd[senv.lo:senv.hi])):
                bad += 1
        if mode in ("fused_async", "fused_async16"):
            # pipelined semantics: after the kernel of step t + lag (no flush, nothing else launched) every rank holds all
            # rows of step t -- forwarded by the next step's kernel and, in the 16-bit format, rebuilt by the one after
            hist = []
            for t in range(10):
                a = torch.from_numpy(random_actions(rng, E)).to(dev)
                senv.step_local(a[senv.lo:senv.hi].contiguous())
                got = senv.wait_pushed()
                fo, _, _ = full.step(a)
                torch.cuda.synchronize()
                dist.barrier()
                hist.append(fo.clone())
                k = len(hist) - 1 - senv.lag
                if k >= 0 and (got is None or not torch.equal(got.view(torch.int32), hist[k].view(torch.int32))):
                    bad += 1
            if not torch.equal(senv.wait_gathered().view(torch.int32), hist[-1].view(torch.int32)):   # flush: the latest step
                bad += 1
            torch.cuda.synchronize()
            dist.barrier()
        if mode in ("fused", "fused_mc", "fused_async", "fused_async16"):
            # the same steps as ONE CUDA graph, replayed twice: the step counting of the fused gather lives on the
            # device, so every replay must wait for / signal the right steps (6 steps = two turns of the 3 buffers)
            acts = [torch.from_numpy(random_actions(rng, E)).to(dev) for _ in range(6)]
            loc = [a[senv.lo:senv.hi].contiguous() for a in acts]
            torch.cuda.synchronize()
            dist.barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for a in loc:
                    senv.step_local(a)
                senv.wait_gathered()
            for rep in range(2):
                g.replay()
                for a in acts:
                    fo, fr, fd = full.step(a)
                torch.cuda.synchronize()
                dist.barrier()
                if not torch.equal(senv.obs_all.view(torch.int32), fo.view(torch.int32)):
                    bad += 1
            bad += senv.env.gather_timeouts
        q.put((rank, bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["fused", "fused_mc", "fused_async", "fused_async16", "collective"])
def test_two_gpu_gather_equals_single_gpu(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    E, steps = 1000, 40        # 500 worlds per rank: ragged last tile, odd tile counts
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, E, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    if mode == "fused_mc" and any(r[1] == "no-multicast" for r in res):
        pytest.skip("symmetric memory has no multicast pointer on this box")
    assert sorted(res) == [(0, 0), (1, 0)], "gathered rows differ from the single-GPU batch: %r" % (res,)
