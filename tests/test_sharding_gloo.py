"""The N>1 path on CPU: world_size-2 gloo, env-id sharding + in-place obs all-gather.

The rank-local stepper here is the oracle (a stand-in with CrowdNavVecEnv's surface);
on the GPU box bench.py runs the same ShardedVecEnv over CrowdNavVecEnv + NCCL.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crowdnav_b200.config import baseline_config
from crowdnav_b200.sharded import ShardedVecEnv, local_config, shard_range
from oracle.oracle import OracleEnv
from parity_util import random_actions


class _OracleStepper:
    """CrowdNavVecEnv's reset()/step() surface over the CPU oracle, writing into obs_out."""

    def __init__(self, cfg, obs_out):
        self.o = OracleEnv(cfg)
        self.obs = obs_out

    def reset(self):
        self.obs.copy_(torch.from_numpy(self.o.reset()))
        return self.obs

    def step(self, actions):
        ob, r, d = self.o.step(actions.numpy())
        self.obs.copy_(torch.from_numpy(ob))
        return self.obs, torch.from_numpy(r.copy()), torch.from_numpy(d.copy())


def test_shard_range_partitions_exactly():
    for E in (1, 7, 64, 4096, 65536):
        for W in (1, 2, 3, 8):
            spans = [shard_range(E, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == E
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_local_config_offsets():
    cfg = baseline_config(3, n_envs=64, env_id_offset=100)
    c1 = local_config(cfg, 1, 4)
    assert c1.n_envs == 16 and c1.env_id_offset == 116


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, E, steps, q, flags=0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = baseline_config(3, n_envs=E)
        cfg.flags |= flags
        if flags & 8:                       # risk_faithful rows cannot leave through the fused gather
            with pytest.raises(ValueError):
                ShardedVecEnv(cfg, lambda c, o: _OracleStepper(c, o), torch.device("cpu"), gather="fused")
        senv = ShardedVecEnv(cfg, lambda c, o: _OracleStepper(c, o), torch.device("cpu"))
        rng = np.random.default_rng(123)
        acts = [random_actions(rng, E) for _ in range(steps)]
        outs = [senv.reset().clone()]
        for a in acts:
            obs_all, r, d = senv.step(torch.from_numpy(a[senv.lo:senv.hi].copy()))
            outs.append(obs_all.clone())
        if rank == 0:
            q.put(torch.stack(outs).numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("flags", [0, 8], ids=["risk_intended", "risk_faithful"])
def test_two_rank_gather_equals_single_process(flags):
    """Same seeds on 1 vs 2 ranks give bit-identical concatenated observations."""
    E, steps, world = 12, 25, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, E, steps, q, flags)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_cfg = baseline_config(3, n_envs=E)
    ref_cfg.flags |= flags
    ref = OracleEnv(ref_cfg)
    rng = np.random.default_rng(123)
    acts = [random_actions(rng, E) for _ in range(steps)]
    want = [ref.reset().copy()]
    for a in acts:
        want.append(ref.step(a)[0].copy())
    assert np.array_equal(got.view(np.uint32), np.stack(want).view(np.uint32))
