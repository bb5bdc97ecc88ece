"""Multi-GPU layout: one process per GPU, worlds sharded by env id.

Every world is independent (no cross-env read anywhere in the reference's Env),
so the path has NO data-path exchange: rank r owns the contiguous block of
global env ids [r * E/G, (r+1) * E/G) and steps it with its own kernel.  RNG
streams are keyed by GLOBAL env id (cn_config.env_id_offset), so results do not
depend on G.  The one collective is what BASELINE.json's north_star asks for:
an all-gather of the observation tensor so every rank (learner replicas) sees
all E rows.  The step kernel writes straight into this rank's slice of the
gather buffer and the all-gather runs in place (NCCL over NVLink / NVSwitch).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist

from .config import CnConfig


def shard_range(n_envs_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous env-id block [lo, hi) of `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(n_envs_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_config(cfg_global: CnConfig, rank: int, world: int) -> CnConfig:
    """This rank's cn_config: its env count and the global id of its env 0."""
    lo, hi = shard_range(cfg_global.n_envs, rank, world)
    cfg = cfg_global.copy()
    cfg.n_envs = hi - lo
    cfg.env_id_offset = cfg_global.env_id_offset + lo
    return cfg


class ShardedVecEnv:
    """E worlds over `world` ranks; step() returns the local (obs, reward, done)
    views and leaves the gathered [E, D] observation in ``self.obs_all``.

    `make_local(cfg_local, obs_slice)` builds the rank's stepper: on a GPU box it
    is ``CrowdNavVecEnv(cfg_local, obs_out=obs_slice)``; the CPU gloo tests pass
    a stand-in with the same reset()/step() surface.
    """

    def __init__(self, cfg_global: CnConfig, make_local: Callable, device: torch.device,
                 group: dist.ProcessGroup | None = None, gather: str = "collective"):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if cfg_global.n_envs % self.world != 0:
            raise ValueError("all_gather_into_tensor needs equal shards: n_envs %% world_size must be 0")
        self.cfg_global = cfg_global.copy()
        self.cfg_local = local_config(cfg_global, self.rank, self.world)
        self.lo, self.hi = shard_range(cfg_global.n_envs, self.rank, self.world)
        self.E, self.D = cfg_global.n_envs, cfg_global.obs_dim
        self.gather_mode = gather if self.world > 1 else "none"
        if self.gather_mode == "fused" and (cfg_global.flags & 8):
            # CN_FLAG_RISK_FAITHFUL: cn_faithful_kernel rewrites the K block after the step kernel has already sent
            # its rows to the peers, so cn_step_gather refuses peers in that mode
            raise ValueError("risk_faithful worlds gather with gather='collective' (ncclAllGather), not 'fused'")
        self._symm = None
        if self.gather_mode == "fused":
            # Gather buffers in symmetric memory: every rank can address every peer's copy, so the step kernel
            # stores its rows into all of them itself (bulk TMA stores over NVLink) -- no collective launch.
            # What is left of the collective is a cross-rank barrier ("every shard of step t has landed"); it runs
            # on a side stream, off the stepping stream's critical path.  THREE buffers rotate: the kernel of
            # step t writes buffer t % 3 on every rank, which a slow peer may still be reading from step t-3; that
            # peer finished reading before it launched step t-2, which is what barrier t-2 certifies -- so the
            # stepping stream only ever waits for a barrier issued two steps earlier.
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            self._bufs, self._symm, self._peers = [], [], []
            for _ in range(3):
                buf = symm_mem.empty((self.E, self.D), dtype=torch.float32, device=device)
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, grp)
                off = self.lo * self.D * 4
                self._bufs.append(buf)
                self._symm.append(hdl)
                self._peers.append([int(p) + off for r, p in enumerate(hdl.buffer_ptrs) if r != self.rank])
            self._cur = 0
            self._comm = torch.cuda.Stream(device=device, priority=-1)   # barrier kernels jump the queue
            self._ready = [None, None, None]                # event: barrier of the step that wrote buffer i is done
            self.obs_all = self._bufs[0]
        else:
            self.obs_all = torch.zeros((self.E, self.D), dtype=torch.float32, device=device)
        self.obs_local = self.obs_all[self.lo:self.hi]          # contiguous row block
        self.env = make_local(self.cfg_local, self.obs_local)
        if self.gather_mode == "fused":
            self.env.set_obs_peers(self._peers[0])

    def gather(self) -> torch.Tensor:
        """Make obs_all complete on every rank.  'collective': an in-place all-gather of the rows (sendbuf =
        recvbuf + rank * count) on the stepping stream.  'fused': the kernels already wrote every peer's copy; the
        cross-rank barrier is enqueued on the side stream -- call wait_gathered() before reading other ranks' rows."""
        if self.gather_mode == "collective":
            dist.all_gather_into_tensor(self.obs_all, self.obs_local, group=self.group)
        elif self.gather_mode == "fused":
            main = torch.cuda.current_stream(self.obs_all.device)
            done = torch.cuda.Event()
            done.record(main)
            self._comm.wait_event(done)
            with torch.cuda.stream(self._comm):
                self._symm[self._cur].barrier(channel=0)
                ev = torch.cuda.Event()
                ev.record(self._comm)
            self._ready[self._cur] = ev
        return self.obs_all

    def wait_gathered(self) -> torch.Tensor:
        """Order the current stream behind the arrival of every rank's rows of the latest step."""
        if self.gather_mode == "fused" and self._ready[self._cur] is not None:
            torch.cuda.current_stream(self.obs_all.device).wait_event(self._ready[self._cur])
        return self.obs_all

    def reset(self) -> torch.Tensor:
        self.env.reset()
        if self.gather_mode == "fused":
            # cn_reset writes the local rows only: publish them once with the collective
            self._symm[self._cur].barrier(channel=0)
            dist.all_gather_into_tensor(self.obs_all, self.obs_local, group=self.group)
            self._symm[self._cur].barrier(channel=0)
            return self.obs_all
        return self.gather()

    def begin_step(self) -> None:
        """Fused mode: rotate to the next gather buffer (see __init__) once its previous readers are done."""
        if self.gather_mode == "fused":
            self._cur = (self._cur + 1) % 3
            guard = self._ready[(self._cur + 1) % 3]         # barrier of two steps ago
            if guard is not None:
                torch.cuda.current_stream(self.obs_all.device).wait_event(guard)
            self.obs_all = self._bufs[self._cur]
            self.obs_local = self.obs_all[self.lo:self.hi]
            self.env.obs = self.obs_local
            self.env.set_obs_peers(self._peers[self._cur])

    def step(self, actions_local: torch.Tensor):
        self.begin_step()
        _, reward, done = self.env.step(actions_local)
        self.gather()
        return self.obs_all, reward, done
