"""Multi-GPU layout: one process per GPU, worlds sharded by env id.

Every world is independent (no cross-env read anywhere in the reference's Env),
so the path has NO data-path exchange: rank r owns the contiguous block of
global env ids [r * E/G, (r+1) * E/G) and steps it with its own kernel.  RNG
streams are keyed by GLOBAL env id (cn_config.env_id_offset), so results do not
depend on G.  The one collective is what BASELINE.json's north_star asks for:
an all-gather of the observation tensor so every rank (learner replicas) sees
all E rows.  The step kernel writes straight into this rank's slice of the
gather buffer and either the all-gather runs in place (NCCL over NVLink /
NVSwitch) or -- the default on GPUs -- the step kernel itself stores its rows
into every peer's buffer and signals them (see ShardedVecEnv).
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


GATHER_MODES = ("collective", "fused", "fused_mc", "fused_async", "fused_async16", "none")


class ShardedVecEnv:
    """E worlds over `world` ranks; step() returns the local (obs, reward, done)
    views and leaves the gathered [E, D] observation in ``self.obs_all``.

    `make_local(cfg_local, obs_slice)` builds the rank's stepper: on a GPU box it
    is ``CrowdNavVecEnv(cfg_local, obs_out=obs_slice)``; the CPU gloo tests pass
    a stand-in with the same reset()/step() surface.

    gather modes (world > 1):

    ``"collective"``    the step kernel writes this rank's slice of the [E, D] buffer, then one in-place
                        ``all_gather_into_tensor`` (ncclAllGather) on the stepping stream.
    ``"fused"``         cn_step_gather_signal: the gather buffers and one array of arrival counters per rank live in
                        symmetric memory (peer-mapped over NVLink); the step kernel stores every tile of rows into all
                        peers' buffers itself and then signals each peer's counter -- data AND synchronisation of the
                        collective are inside the one kernel, nothing else is launched per step, and a sequence of
                        steps is graph-capturable.  ``wait_gathered()`` enqueues a one-warp kernel that holds the
                        stream until all rows of the latest step have arrived.  THREE buffers rotate: step t writes
                        buffer t % 3 on every rank, which a peer may have been reading as step t-3's result; that
                        peer finished reading before it launched step t-2 (stream order), and its step-t-2 CTAs
                        signalling THIS rank is what the kernel of step t waits for before its first remote store.
                        All step counting happens on the device (a rank's own counter slot), so a CUDA graph of
                        3k such steps can be replayed any number of times.
    ``"fused_mc"``      the same with NVSwitch multicast: one ``multimem.st`` per 16 bytes reaches every rank's buffer
                        (egress 1x instead of (world-1)x), signal by ``multimem.red``.  Needs multicast-capable
                        symmetric memory; raises if the handle has no multicast pointer.
    ``"fused_async"``   cn_step_gather_async, the PIPELINED fused gather: the kernel of step t+1 forwards the rows of
                        step t to the peers at its start (bulk load into a staging tile, bulk stores to every peer), so
                        the NVLink transfer runs under the next step's compute instead of behind the step that
                        produced the rows.  ``step_local()`` therefore leaves the rows of the step it has just launched
                        un-forwarded; ``wait_gathered()`` flushes them with a push-only launch (cn_gather_flush) and
                        waits, ``wait_pushed()`` only waits for what the step kernels have forwarded so far (all rows
                        of the step BEFORE the latest one, in ``obs_all_prev``).  Rows of step t must be consumed
                        before step t+2 is launched.  Local handles are created with CN_FLAG_GATHER_STAGE.
    ``"fused_async16"`` the same pipeline with a 16-bit wire format: every value of a row is a whole number of thousandths
                        (that is how the reference rounds the row, ENV:1042), so the step kernel also writes its rows as
                        int16 thousandths, the next step's kernel forwards THOSE (half the NVLink bytes), and the kernel
                        after that rebuilds the identical fp32 rows on the receiver while its state tile is loading: ONE
                        launch per step computes step t+1, forwards step t and completes the gather of step t-1
                        (``lag`` = 2; four buffers rotate; at the end of its step a CTA checks that the peers have
                        finished their previous kernel -- which certifies the delivery -- and rebuilds its share of rows
                        while its own pushes drain).
                        ``wait_gathered()`` = flush + wait + cn_gather_decode16 of whatever is still on the wire.  Values
                        must stay below 32.768 in magnitude (rooms within +-30 m); a value that does not fit is counted
                        (``env.gather_timeouts``).
    """

    def __init__(self, cfg_global: CnConfig, make_local: Callable, device: torch.device,
                 group: dist.ProcessGroup | None = None, gather: str = "collective"):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if cfg_global.n_envs % self.world != 0:
            raise ValueError("all_gather_into_tensor needs equal shards: n_envs %% world_size must be 0")
        if gather not in GATHER_MODES:
            raise ValueError("unknown gather mode %r" % gather)
        self.cfg_global = cfg_global.copy()
        self.gather_mode = gather if self.world > 1 else "none"
        self.fused = self.gather_mode.startswith("fused")
        self.cfg_local = local_config(cfg_global, self.rank, self.world)
        self.async_mode = self.gather_mode in ("fused_async", "fused_async16")
        self.wire16 = self.gather_mode == "fused_async16"
        if self.async_mode:
            self.cfg_local.flags |= 16 | (32 if self.wire16 else 0)      # CN_FLAG_GATHER_STAGE (| CN_FLAG_GATHER_WIRE16)
        self.lo, self.hi = shard_range(cfg_global.n_envs, self.rank, self.world)
        self.E, self.D = cfg_global.n_envs, cfg_global.obs_dim
        if self.fused and (cfg_global.flags & 8):
            # CN_FLAG_RISK_FAITHFUL: cn_faithful_kernel rewrites the K block after the step kernel has already sent
            # its rows to the peers, so the fused entry points refuse that mode
            raise ValueError("risk_faithful worlds gather with gather='collective' (ncclAllGather), not 'fused'")
        self._symm = None
        self._pending = False                    # pipelined modes: the latest step's rows have not been forwarded yet
        self._inflight = None                    # fused_async16: buffer forwarded by the last kernel, not yet rebuilt
        self._complete = None                    # newest buffer whose rows are complete on this rank
        self._to_decode = []
        if self.fused:
            import ctypes as C
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            self._bufs, self._symm, self._peers, self._mc = [], [], [], []
            # peers in ring order starting at rank + 1: at any moment the ranks' stores are headed for DIFFERENT
            # GPUs (ascending order from 0 would aim every rank at GPU 0 first, then GPU 1, ...)
            order = [(self.rank + 1 + k) % self.world for k in range(self.world - 1)]
            off = self.lo * self.D * 4
            self._wire, self._wsymm = [], []
            # gather buffers in rotation: three cover a lag of one step; the 16-bit pipeline completes a step two launches
            # later, and the buffer must not be reused (this rank's own rows!) before that step has been consumed: four
            self._R = 4 if self.wire16 else 3
            for _ in range(self._R):
                buf = symm_mem.empty((self.E, self.D), dtype=torch.float32, device=device)
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, grp)
                self._bufs.append(buf)
                self._symm.append(hdl)
                if self.wire16:
                    # the peers write int16 rows into the wire buffers; the fp32 buffers stay local (decode target)
                    wbuf = symm_mem.empty((self.E, self.D), dtype=torch.int16, device=device)
                    wbuf.zero_()
                    whdl = symm_mem.rendezvous(wbuf, grp)
                    self._wire.append(wbuf)
                    self._wsymm.append(whdl)
                    ptrs = [int(whdl.buffer_ptrs[r]) + off // 2 for r in order]
                    hdl = whdl
                else:
                    ptrs = [int(hdl.buffer_ptrs[r]) + off for r in order]
                self._peers.append((C.c_void_p * len(ptrs))(*ptrs))
                mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                self._mc.append(mc + off if mc else 0)
            # arrival counters: one 64-bit slot per SOURCE rank (padded to 32 slots)
            self._arrive = symm_mem.empty((32,), dtype=torch.int64, device=device)
            self._arrive.zero_()
            ah = symm_mem.rendezvous(self._arrive, grp)
            self._arrive_hdl = ah
            aptrs = [int(ah.buffer_ptrs[r]) + 8 * self.rank for r in order]
            self._arrive_peers = (C.c_void_p * len(aptrs))(*aptrs)
            amc = int(getattr(ah, "multicast_ptr", 0) or 0)
            self._arrive_mc = amc + 8 * self.rank if amc else 0
            if self.gather_mode == "fused_mc" and not (amc and all(self._mc)):
                raise RuntimeError("gather='fused_mc' needs multicast-capable symmetric memory (multicast_ptr is 0)")
            self.multicast_available = bool(amc and all(self._mc))
            torch.cuda.synchronize(device)
            ah.barrier(channel=0)                # everyone's counters are zero before anyone signals
            self._cur = 0
            self.obs_all = self._bufs[0]
            self.obs_all_prev = self._bufs[0]
        else:
            self.obs_all = torch.zeros((self.E, self.D), dtype=torch.float32, device=device)
        self.obs_local = self.obs_all[self.lo:self.hi]          # contiguous row block
        self.env = make_local(self.cfg_local, self.obs_local)

    def step_local(self, actions_local: torch.Tensor, env=None):
        """Rotate the gather buffer and launch this rank's step (with the fused gather where enabled) on the current
        stream.  `env`: a replica of the local shard (same config) to step instead of self.env -- bench.py rotates
        several so that their state is cold in L2.  Returns the stepper's (obs_local, reward, done)."""
        env = self.env if env is None else env
        if self.fused:
            prev = self._cur
            self._cur = (self._cur + 1) % self._R
            self.obs_all_prev = self._bufs[prev]
            self.obs_all = self._bufs[self._cur]
            self.obs_local = self.obs_all[self.lo:self.hi]
            env.obs = self.obs_local
            if self.wire16:
                # one launch: compute this step (fp32 rows + int16 copy into wire[cur]), forward the previous step's int16
                # rows, rebuild the fp32 rows the peers' previous kernels delivered (the step before that)
                wire_out = self._wire[self._cur][self.lo:self.hi].data_ptr()
                if self._pending:
                    dec = self._inflight
                    env.step_gather_async(actions_local, self._wire[prev][self.lo:self.hi].data_ptr(), self._peers[prev],
                                          self._arrive_peers, self.world - 1, self._arrive.data_ptr(), self.world, self.rank, 2,
                                          wire_out=wire_out,
                                          dec_wire=self._wire[dec].data_ptr() if dec is not None else 0,
                                          dec_obs=self._bufs[dec].data_ptr() if dec is not None else 0)
                    if dec is not None:
                        self._complete = dec
                    self._inflight = prev
                else:                            # nothing to forward (first step, or the rows were flushed): encode only
                    env.step_gather_async(actions_local, 0, None, None, 0, 0, self.world, self.rank, 2, wire_out=wire_out)
                self._pending = True
            elif self.async_mode:
                if self._pending:
                    env.step_gather_async(actions_local, self._bufs[prev][self.lo:self.hi].data_ptr(), self._peers[prev],
                                          self._arrive_peers, self.world - 1, self._arrive.data_ptr(), self.world, self.rank, 2)
                else:
                    env.step(actions_local)      # nothing to forward (first step, or the rows were flushed)
                self._complete = prev if self._pending else self._complete
                self._pending = True
            else:
                mc = self.gather_mode == "fused_mc"
                env.step_gather_signal(actions_local, self._peers[self._cur], self._arrive_peers, self.world - 1,
                                       self._mc[self._cur] if mc else 0, self._arrive_mc if mc else 0,
                                       self._arrive.data_ptr(), self.world, self.rank, 2)
            return env.obs, env.reward, env.done
        if env is not self.env:
            env.obs = self.obs_local
        return env.step(actions_local)

    def gather(self) -> torch.Tensor:
        """Make obs_all complete on every rank.  'collective': an in-place all-gather of the rows (sendbuf =
        recvbuf + rank * count) on the stepping stream.  Fused modes: nothing to launch here -- the kernels write
        every peer's copy and signal; call wait_gathered() before reading other ranks' rows."""
        if self.gather_mode == "collective":
            dist.all_gather_into_tensor(self.obs_all, self.obs_local, group=self.group)
        return self.obs_all

    def flush_gather(self) -> None:
        """Pipelined modes: forward the rows of the latest step now (push-only launch) instead of with the next step."""
        if self.async_mode and self._pending:
            src = self._wire[self._cur][self.lo:self.hi].data_ptr() if self.wire16 else self.obs_local.data_ptr()
            self.env.gather_flush(src, self._peers[self._cur], self._arrive_peers, self.world - 1,
                                  self._arrive.data_ptr(), self.world, self.rank, 2, wire16=self.wire16)
            self._pending = False
            if self.wire16:
                self._to_decode = [b for b in (self._inflight, self._cur) if b is not None]
                self._inflight = None

    def wait_pushed(self) -> torch.Tensor | None:
        """The newest gather buffer that is COMPLETE on this rank without flushing anything: 'fused_async' waits for what
        the step kernels have forwarded and returns the step before the latest one; 'fused_async16' returns the step
        two before the latest (rebuilt inside the latest step's kernel, nothing to wait for); the non-pipelined fused
        modes wait for the latest step itself.  None while the pipeline is still filling."""
        if not self.fused:
            return self.obs_all
        if self.wire16:
            return self._bufs[self._complete] if self._complete is not None else None
        self.env.gather_wait(self._arrive.data_ptr(), self.world, self.rank)
        if self.async_mode:
            return self._bufs[self._complete] if self._complete is not None else None
        return self.obs_all

    @property
    def lag(self) -> int:
        """Steps between a step's launch and the launch whose completion makes its rows complete on every rank."""
        return {"fused_async": 1, "fused_async16": 2}.get(self.gather_mode, 0)

    def wait_gathered(self) -> torch.Tensor:
        """Order the current stream behind the arrival of every rank's rows of the LATEST step."""
        if self.fused:
            self.flush_gather()
            self.env.gather_wait(self._arrive.data_ptr(), self.world, self.rank)
            for b in getattr(self, "_to_decode", []):
                self.env.gather_decode16(self._wire[b].data_ptr(), self._bufs[b].data_ptr(), self.lo, self.hi, self.E)
            self._to_decode = []
            self._complete = self._cur
        return self.obs_all

    def reset(self) -> torch.Tensor:
        self.env.obs = self.obs_local
        self.env.reset()
        if self.fused:
            # cn_reset writes the local rows only: publish them once with the collective
            self._symm[self._cur].barrier(channel=0)
            dist.all_gather_into_tensor(self.obs_all, self.obs_local, group=self.group)
            self._symm[self._cur].barrier(channel=0)
            self._pending, self._inflight, self._to_decode = False, None, []
            self._complete = self._cur
            return self.obs_all
        return self.gather()

    def step_host(self, actions_local):
        """The sharded step through HOST buffers (bench.py's e2e leg for N > 1): pinned H2D of this rank's actions, the
        step (+ gather), arrival of every rank's rows certified, D2H of the local obs / reward / done, stream sync."""
        env = self.env
        env._host_buffers()
        env._h_act.numpy()[...] = actions_local
        env._d_act.copy_(env._h_act, non_blocking=True)
        _, reward, done = self.step_local(env._d_act)
        self.gather()
        self.wait_gathered()
        env._h_obs.copy_(self.obs_local, non_blocking=True)
        env._h_rew.copy_(reward, non_blocking=True)
        env._h_done.copy_(done, non_blocking=True)
        torch.cuda.current_stream(self.obs_all.device).synchronize()
        return env._h_obs.numpy(), env._h_rew.numpy(), env._h_done.numpy()

    def step(self, actions_local: torch.Tensor):
        """One strict step: launches it and the gather ('collective'); in the fused modes call wait_gathered() before
        reading other ranks' rows of obs_all."""
        _, reward, done = self.step_local(actions_local)
        self.gather()
        return self.obs_all, reward, done
