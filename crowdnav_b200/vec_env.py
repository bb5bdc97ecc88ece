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

    # -- lifecycle -------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.cn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- the path ---------------------------------------------------------------
    def reset(self, mask: torch.Tensor | None = None) -> torch.Tensor:
        """Env.reset for the masked worlds (all if mask is None)."""
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        _lib.check(self._L.cn_reset(self._h, mp, C.c_void_p(self.obs.data_ptr()), self._stream()), "cn_reset")
        return self.obs

    def step(self, actions: torch.Tensor):
        """One control period.  actions: float32 [E, 2] (v, w) on this device."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous() \
                or actions.numel() != 2 * self.E:
            raise ValueError("actions must be a contiguous float32 [E, 2] tensor on %s" % self.device)
        if self._peer_ptrs is not None:
            _lib.check(self._L.cn_step_gather(self._h, C.c_void_p(actions.data_ptr()), C.c_void_p(self.obs.data_ptr()),
                                              self._peer_ptrs, len(self._peer_ptrs), C.c_void_p(self.reward.data_ptr()),
                                              C.c_void_p(self.done.data_ptr()), self._stream()), "cn_step_gather")
        else:
            _lib.check(self._L.cn_step(self._h, C.c_void_p(actions.data_ptr()), C.c_void_p(self.obs.data_ptr()),
                                       C.c_void_p(self.reward.data_ptr()), C.c_void_p(self.done.data_ptr()),
                                       self._stream()), "cn_step")
        return self.obs, self.reward, self.done

    def set_obs_peers(self, ptrs) -> None:
        """Fuse the observation all-gather into the step kernel: `ptrs` are peer-mapped device addresses of this
        rank's row block inside each PEER's [E_total, D] gather buffer (see ShardedVecEnv, gather='fused')."""
        ptrs = [int(p) for p in ptrs]
        if len(ptrs) > 8:
            raise ValueError("at most 8 peers")
        self._peer_ptrs = (C.c_void_p * len(ptrs))(*ptrs) if ptrs else None

    def step_host(self, actions: np.ndarray):
        """Same step through HOST buffers: pinned H2D of the actions, the kernel,
        D2H of obs / reward / done, one stream synchronise.  This is the call the
        single-env `Env` shim and bench.py's e2e leg make."""
        if self._h_act is None:
            self._h_act = torch.zeros((self.E, 2), dtype=torch.float32).pin_memory()
            self._h_obs = torch.zeros((self.E, self.D), dtype=torch.float32).pin_memory()
            self._h_rew = torch.zeros(self.E, dtype=torch.float32).pin_memory()
            self._h_done = torch.zeros(self.E, dtype=torch.uint8).pin_memory()
            self._d_act = torch.zeros((self.E, 2), dtype=torch.float32, device=self.device)
        self._h_act.numpy()[...] = np.asarray(actions, dtype=np.float32).reshape(self.E, 2)
        self._d_act.copy_(self._h_act, non_blocking=True)
        self.step(self._d_act)
        self._h_obs.copy_(self.obs, non_blocking=True)
        self._h_rew.copy_(self.reward, non_blocking=True)
        self._h_done.copy_(self.done, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._h_obs.numpy(), self._h_rew.numpy(), self._h_done.numpy()

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.E * 2 * 4

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.E * (self.D * 4 + 4 + 1)

    # -- episode bookkeeping (ENV:1265-1283) -----------------------------------------
    def counters(self) -> torch.Tensor:
        """int32 [E, 4]: success, ego violations, social violations, obstacle-present steps."""
        _lib.check(self._L.cn_get_counters(self._h, C.c_void_p(self._counters.data_ptr()), self._stream()),
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
