"""Batched host API over the C ABI: E worlds per call, tensors stay on the GPU.

PyTorch is plumbing here (device memory, streams); every env computation is
the CUDA library's.  ``CrowdNavVecEnv`` is what a vectorised rollout driver
uses; ``crowdnav_b200.env.Env`` is the reference's single-env duck type on top
of it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .config import CnConfig


class CrowdNavVecEnv:
    """E independent crowd-navigation worlds on one GPU.

    step()/reset() enqueue one kernel on torch's current stream and return
    device tensors owned by the env (overwritten by the next call).
    """

    def __init__(self, cfg: CnConfig, device: int | None = None, obs_out: torch.Tensor | None = None):
        if not torch.cuda.is_available():
            raise _lib.CrowdNavError("CrowdNavVecEnv needs a CUDA device; there is no CPU fallback")
        self._L = _lib.load()
        self.cfg = cfg.copy()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.E, self.D, self.NR = cfg.n_envs, cfg.obs_dim, cfg.n_samples - 1
        h = C.c_void_p()
        _lib.check(self._L.cn_create(C.byref(self.cfg), self.device_index, C.byref(h)), "cn_create")
        self._h = h
        with torch.cuda.device(self.device):
            if obs_out is not None:
                # e.g. this rank's slice of an NCCL all-gather buffer
                if obs_out.shape != (self.E, self.D) or obs_out.dtype != torch.float32 or not obs_out.is_contiguous():
                    raise ValueError("obs_out must be a contiguous float32 [E, D] tensor")
                self.obs = obs_out
            else:
                self.obs = torch.zeros((self.E, self.D), dtype=torch.float32, device=self.device)
            self.reward = torch.zeros(self.E, dtype=torch.float32, device=self.device)
            self.done = torch.zeros(self.E, dtype=torch.uint8, device=self.device)
            self._counters = torch.zeros((self.E, 4), dtype=torch.int32, device=self.device)
        self._dbg_ranges = None
        self._dbg_hid = None
        self._peer_ptrs = None          # fused all-gather targets (set_obs_peers)
        # pinned staging for the host-buffer path
        self._h_act = self._h_obs = self._h_rew = self._h_done = None
        self._pipe = None               # step_host_pipelined state
        self._graphs = []

    # -- lifecycle -------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            for g in getattr(self, "_graphs", []):
                g.close()
            self._graphs = []
            self._L.cn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- the path ---------------------------------------------------------------
    def reset(self, mask: torch.Tensor | None = None) -> torch.Tensor:
        """Env.reset for the masked worlds (all if mask is None)."""
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        _lib.check(self._L.cn_reset(self._h, mp, self.obs.data_ptr(), self._stream()), "cn_reset")
        return self.obs

    def step(self, actions: torch.Tensor):
        """One control period.  actions: float32 [E, 2] (v, w) on this device."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous() \
                or actions.numel() != 2 * self.E:
            raise ValueError("actions must be a contiguous float32 [E, 2] tensor on %s" % self.device)
        # (plain ints for the pointer arguments: ctypes converts them, no c_void_p objects per call)
        if self._peer_ptrs is not None:
            rc = self._L.cn_step_gather(self._h, actions.data_ptr(), self.obs.data_ptr(), self._peer_ptrs,
                                        len(self._peer_ptrs), self.reward.data_ptr(), self.done.data_ptr(), self._stream())
        else:
            rc = self._L.cn_step(self._h, actions.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                                 self.done.data_ptr(), self._stream())
        if rc != 0:
            _lib.check(rc, "cn_step")
        return self.obs, self.reward, self.done

    def step_n(self, actions: torch.Tensor, reward_out: torch.Tensor | None = None, done_out: torch.Tensor | None = None):
        """n control periods enqueued by ONE library call (cn_step_n): actions float32 [n, E, 2] (or [E, 2] with
        `n` taken from reward_out: the same batch repeated -- action repeat / frame skip).  Returns (obs of the last
        step, reward [n, E], done [n, E])."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            raise ValueError("actions must be a contiguous float32 tensor on %s" % self.device)
        if actions.dim() == 3:
            n, stride = actions.shape[0], 2 * self.E
            if tuple(actions.shape[1:]) != (self.E, 2):
                raise ValueError("actions must be [n, E, 2]")
        else:
            if reward_out is None or actions.numel() != 2 * self.E:
                raise ValueError("[E, 2] actions need reward_out [n, E] to give the repeat count")
            n, stride = reward_out.shape[0], 0
        if reward_out is None:
            reward_out = torch.empty((n, self.E), dtype=torch.float32, device=self.device)
        if done_out is None:
            done_out = torch.empty((n, self.E), dtype=torch.uint8, device=self.device)
        if tuple(reward_out.shape) != (n, self.E) or tuple(done_out.shape) != (n, self.E) \
                or reward_out.dtype != torch.float32 or done_out.dtype != torch.uint8 \
                or not reward_out.is_contiguous() or not done_out.is_contiguous():
            raise ValueError("reward_out / done_out must be contiguous [n, E] float32 / uint8")
        _lib.check(self._L.cn_step_n(self._h, n, actions.data_ptr(), stride, self.obs.data_ptr(), reward_out.data_ptr(),
                                     done_out.data_ptr(), self.E, self._stream()), "cn_step_n")
        return self.obs, reward_out, done_out

    def make_graph(self, actions: torch.Tensor) -> "StepGraph":
        """The n steps of `actions` [n, E, 2] captured once as a CUDA graph owned by the library (cn_graph_create);
        `graph.launch()` replays them with one driver call.  The tensors are kept alive by the returned object."""
        g = StepGraph(self, actions)
        self._graphs.append(g)
        return g

    def step_gather_signal(self, actions: torch.Tensor, peer_obs, peer_arrive, n_peers: int, obs_mc: int, arrive_mc: int,
                           arrive_local: int, n_ranks: int, rank: int, wait_back: int = 2):
        """cn_step_gather_signal (see include/crowdnav.h): the step with the observation all-gather -- data and
        signalling -- fused into the kernel.  peer_obs / peer_arrive are ctypes arrays of peer-mapped addresses."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous() \
                or actions.numel() != 2 * self.E:
            raise ValueError("actions must be a contiguous float32 [E, 2] tensor on %s" % self.device)
        rc = self._L.cn_step_gather_signal(self._h, actions.data_ptr(), self.obs.data_ptr(), peer_obs, peer_arrive, n_peers,
                                           obs_mc or None, arrive_mc or None, arrive_local, n_ranks, rank, wait_back,
                                           self.reward.data_ptr(), self.done.data_ptr(), self._stream())
        if rc != 0:
            _lib.check(rc, "cn_step_gather_signal")
        return self.obs, self.reward, self.done

    def step_gather_async(self, actions: torch.Tensor, push_src: int, push_peers, peer_arrive, n_peers: int,
                          arrive_local: int, n_ranks: int, rank: int, wait_back: int = 2, wire_out: int = 0,
                          dec_wire: int = 0, dec_obs: int = 0):
        """cn_step_gather_async (see include/crowdnav.h): this step's kernel forwards the PREVIOUS step's rows
        (push_src) to the peers under its own compute; with wire_out (this rank's row block of its own int16 wire
        buffer) the rows travel as int16 thousandths; with dec_wire / dec_obs (wait_back = 1) the same kernel also rebuilds
        the fp32 rows the peers' previous kernels delivered.  Handle created with CN_FLAG_GATHER_STAGE."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous() \
                or actions.numel() != 2 * self.E:
            raise ValueError("actions must be a contiguous float32 [E, 2] tensor on %s" % self.device)
        rc = self._L.cn_step_gather_async(self._h, actions.data_ptr(), self.obs.data_ptr(), wire_out or None, push_src or None,
                                          push_peers, peer_arrive, n_peers, arrive_local or None, n_ranks, rank, wait_back,
                                          dec_wire or None, dec_obs or None,
                                          self.reward.data_ptr(), self.done.data_ptr(), self._stream())
        if rc != 0:
            _lib.check(rc, "cn_step_gather_async")
        return self.obs, self.reward, self.done

    def gather_flush(self, push_src: int, push_peers, peer_arrive, n_peers: int, arrive_local: int, n_ranks: int,
                     rank: int, wait_back: int = 2, wire16: bool = False) -> None:
        _lib.check(self._L.cn_gather_flush(self._h, 1 if wire16 else 0, push_src, push_peers, peer_arrive, n_peers, arrive_local,
                                           n_ranks, rank, wait_back, self._stream()), "cn_gather_flush")

    def gather_decode16(self, wire: int, obs_all: int, row_lo: int, row_hi: int, rows_total: int) -> None:
        _lib.check(self._L.cn_gather_decode16(self._h, wire, obs_all, row_lo, row_hi, rows_total, self._stream()),
                   "cn_gather_decode16")

    def gather_wait(self, arrive_local: int, n_ranks: int, rank: int) -> None:
        _lib.check(self._L.cn_gather_wait(self._h, arrive_local, n_ranks, rank, self._stream()), "cn_gather_wait")

    @property
    def gather_timeouts(self) -> int:
        out = C.c_uint32(0)
        _lib.check(self._L.cn_gather_timeouts(self._h, C.byref(out), self._stream()), "cn_gather_timeouts")
        return int(out.value)

    @property
    def kernel_ctas(self) -> int:
        return int(self._L.cn_kernel_ctas(self._h))

    def set_obs_peers(self, ptrs) -> None:
        """Fuse the observation all-gather into the step kernel: `ptrs` are peer-mapped device addresses of this
        rank's row block inside each PEER's [E_total, D] gather buffer (see ShardedVecEnv, gather='fused')."""
        ptrs = [int(p) for p in ptrs]
        if len(ptrs) > 8:
            raise ValueError("at most 8 peers")
        self._peer_ptrs = (C.c_void_p * len(ptrs))(*ptrs) if ptrs else None

    def _host_buffers(self):
        if self._h_act is None:
            self._h_act = torch.zeros((self.E, 2), dtype=torch.float32).pin_memory()
            self._h_obs = torch.zeros((self.E, self.D), dtype=torch.float32).pin_memory()
            self._h_rew = torch.zeros(self.E, dtype=torch.float32).pin_memory()
            self._h_done = torch.zeros(self.E, dtype=torch.uint8).pin_memory()
            self._d_act = torch.zeros((self.E, 2), dtype=torch.float32, device=self.device)

    def step_host(self, actions: np.ndarray, mode: str = "copy"):
        """Same step through HOST buffers; returns numpy views of pinned host memory (overwritten by the next call).
        This is the call the single-env `Env` shim and bench.py's e2e leg make.

        mode="copy":   pinned H2D of the actions, the kernel, three D2H copies (obs / reward / done) by the copy
                       engine, one stream synchronise.
        mode="mapped": the kernel itself writes obs / reward / done into the pinned host buffers (they are
                       device-addressable under unified addressing): the tile's rows leave the SM by bulk store
                       straight over PCIe, no separate D2H copies and no device-side staging of the rows.  Same bytes
                       over the bus, same results; not available for fused-gather / risk_faithful handles."""
        self._host_buffers()
        self._h_act.numpy()[...] = np.asarray(actions, dtype=np.float32).reshape(self.E, 2)
        self._d_act.copy_(self._h_act, non_blocking=True)
        if mode == "mapped":
            if self._peer_ptrs is not None or (self.cfg.flags & 8):
                raise ValueError("mode='mapped' needs a plain handle (no fused gather, no risk_faithful block)")
            _lib.check(self._L.cn_step(self._h, self._d_act.data_ptr(), self._h_obs.data_ptr(), self._h_rew.data_ptr(),
                                       self._h_done.data_ptr(), self._stream()), "cn_step")
        elif mode == "copy":
            self.step(self._d_act)
            self._h_obs.copy_(self.obs, non_blocking=True)
            self._h_rew.copy_(self.reward, non_blocking=True)
            self._h_done.copy_(self.done, non_blocking=True)
        else:
            raise ValueError("mode must be 'copy' or 'mapped'")
        torch.cuda.current_stream(self.device).synchronize()
        return self._h_obs.numpy(), self._h_rew.numpy(), self._h_done.numpy()

    def step_host_pipelined(self, actions: np.ndarray):
        """Double-buffered host-buffer step for off-policy rollouts that can act on ONE-STEP-OLD observations:
        call k enqueues H2D(actions k) + kernel k on the stepping stream and the D2H of its results on a copy stream,
        then returns the results of step k-1 (None on the first call) -- the host never waits for the step it has just
        issued, and the D2H of step k-1 overlaps kernel k.  flush_host_pipeline() returns the last step's results.
        Opt-in: `step_host` (strict, results of the step just issued) stays the default."""
        if self._pipe is None:
            pin = lambda *shape, dt=torch.float32: torch.zeros(shape, dtype=dt).pin_memory()
            self._pipe = {
                "k": 0,
                "h_act": [pin(self.E, 2), pin(self.E, 2)],
                "d_act": [torch.zeros((self.E, 2), dtype=torch.float32, device=self.device) for _ in range(2)],
                "d_obs": [torch.zeros((self.E, self.D), dtype=torch.float32, device=self.device) for _ in range(2)],
                "d_rew": [torch.zeros(self.E, dtype=torch.float32, device=self.device) for _ in range(2)],
                "d_done": [torch.zeros(self.E, dtype=torch.uint8, device=self.device) for _ in range(2)],
                "h_obs": [pin(self.E, self.D), pin(self.E, self.D)],
                "h_rew": [pin(self.E), pin(self.E)],
                "h_done": [pin(self.E, dt=torch.uint8), pin(self.E, dt=torch.uint8)],
                "copied": [None, None],          # event: D2H of the step that used slot i is complete
                "copy_stream": torch.cuda.Stream(device=self.device),
            }
        p = self._pipe
        k, i = p["k"], p["k"] & 1
        main = torch.cuda.current_stream(self.device)
        if p["copied"][i] is not None:
            p["copied"][i].synchronize()        # the host may reuse slot i's pinned action buffer / read its results
        p["h_act"][i].numpy()[...] = np.asarray(actions, dtype=np.float32).reshape(self.E, 2)
        p["d_act"][i].copy_(p["h_act"][i], non_blocking=True)
        _lib.check(self._L.cn_step(self._h, p["d_act"][i].data_ptr(), p["d_obs"][i].data_ptr(), p["d_rew"][i].data_ptr(),
                                   p["d_done"][i].data_ptr(), main.cuda_stream), "cn_step")
        stepped = torch.cuda.Event()
        stepped.record(main)
        cs = p["copy_stream"]
        cs.wait_event(stepped)
        with torch.cuda.stream(cs):
            p["h_obs"][i].copy_(p["d_obs"][i], non_blocking=True)
            p["h_rew"][i].copy_(p["d_rew"][i], non_blocking=True)
            p["h_done"][i].copy_(p["d_done"][i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        p["copied"][i] = ev
        # the kernel of call k+1 overwrites slot i^1's device buffers: it must wait for their D2H (issued by call k-1)
        j = i ^ 1
        p["k"] = k + 1
        if p["copied"][j] is None:
            return None
        main.wait_event(p["copied"][j])
        p["copied"][j].synchronize()
        return p["h_obs"][j].numpy(), p["h_rew"][j].numpy(), p["h_done"][j].numpy()

    def flush_host_pipeline(self):
        """Results of the last step issued by step_host_pipelined."""
        p = self._pipe
        if p is None or p["k"] == 0:
            return None
        i = (p["k"] - 1) & 1
        p["copied"][i].synchronize()
        return p["h_obs"][i].numpy(), p["h_rew"][i].numpy(), p["h_done"][i].numpy()

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.E * 2 * 4

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.E * (self.D * 4 + 4 + 1)

    # -- episode bookkeeping (ENV:1265-1283) -----------------------------------------
    def counters(self) -> torch.Tensor:
        """int32 [E, 4]: success, ego violations, social violations, obstacle-present steps."""
        _lib.check(self._L.cn_get_counters(self._h, self._counters.data_ptr(), self._stream()),
                   "cn_get_counters")
        return self._counters

    def clear_done(self, mask: torch.Tensor | None = None) -> None:
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        _lib.check(self._L.cn_clear_done(self._h, mp, self._stream()), "cn_clear_done")

    # -- snapshots / debug --------------------------------------------------------
    def get_state_blob(self) -> np.ndarray:
        n = self._L.cn_blob_bytes(C.byref(self.cfg))
        out = np.zeros(n // 4, dtype=np.uint32)
        _lib.check(self._L.cn_get_blob(self._h, out.ctypes.data_as(C.c_void_p), n, self._stream()), "cn_get_blob")
        return out

    def set_state_blob(self, blob: np.ndarray) -> None:
        blob = np.ascontiguousarray(blob, dtype=np.uint32)
        _lib.check(self._L.cn_set_blob(self._h, blob.ctypes.data_as(C.c_void_p), blob.nbytes, self._stream()),
                   "cn_set_blob")

    def enable_debug_taps(self) -> None:
        """Also store pre-rounding ranges and hit ids each step (tests only)."""
        self._dbg_ranges = torch.zeros((self.E, self.NR), dtype=torch.float32, device=self.device)
        self._dbg_hid = torch.zeros((self.E, self.NR), dtype=torch.uint8, device=self.device)
        _lib.check(self._L.cn_set_debug_taps(self._h, C.c_void_p(self._dbg_ranges.data_ptr()),
                                             C.c_void_p(self._dbg_hid.data_ptr())), "cn_set_debug_taps")

    @property
    def debug_ranges(self) -> torch.Tensor:
        return self._dbg_ranges

    @property
    def debug_hit_ids(self) -> torch.Tensor:
        return self._dbg_hid

    @property
    def launch_count(self) -> int:
        return int(self._L.cn_launch_count(self._h))

    @property
    def kernel_name(self) -> str:
        """Step-kernel variant behind this handle (see cn_kernel_name in include/crowdnav.h)."""
        return self._L.cn_kernel_name(self._h).decode()

    @property
    def kernel_tile(self) -> int:
        return int(self._L.cn_kernel_tile(self._h))


class StepGraph:
    """n env steps captured once as a CUDA graph inside the library (cn_graph_create / cn_graph_launch)."""

    def __init__(self, env: CrowdNavVecEnv, actions: torch.Tensor):
        if actions.device != env.device or actions.dtype != torch.float32 or not actions.is_contiguous() \
                or actions.dim() != 3 or tuple(actions.shape[1:]) != (env.E, 2):
            raise ValueError("actions must be a contiguous float32 [n, E, 2] tensor on %s" % env.device)
        self.env, self.actions, self.n = env, actions, actions.shape[0]
        self.reward = torch.zeros((self.n, env.E), dtype=torch.float32, device=env.device)
        self.done = torch.zeros((self.n, env.E), dtype=torch.uint8, device=env.device)
        g = C.c_void_p()
        torch.cuda.synchronize(env.device)
        _lib.check(env._L.cn_graph_create(env._h, self.n, actions.data_ptr(), 2 * env.E, env.obs.data_ptr(),
                                          self.reward.data_ptr(), self.done.data_ptr(), env.E, C.byref(g)), "cn_graph_create")
        self._g = g

    def launch(self):
        _lib.check(self.env._L.cn_graph_launch(self._g, self.env._stream()), "cn_graph_launch")
        return self.env.obs, self.reward, self.done

    def close(self):
        if self._g:
            self.env._L.cn_graph_destroy(self._g)
            self._g = None
