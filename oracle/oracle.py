"""ctypes front-end of the CPU oracle (oracle/cn_oracle.c).

TEST INFRASTRUCTURE: import only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from crowdnav_b200.config import CnConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcn_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle in-tree (gcc, a second or two)."""
    srcs = [os.path.join(_HERE, "cn_oracle.c"), os.path.join(_HERE, "cn_oracle_faithful.c"),
            os.path.join(_HERE, "..", "crowdnav_b200", "csrc", "cn_math64.h"),
            os.path.join(_HERE, "..", "crowdnav_b200", "csrc", "cn_faithful_state.h"),
            os.path.join(_HERE, "..", "crowdnav_b200", "csrc", "cn_math.h"),
            os.path.join(_HERE, "..", "crowdnav_b200", "csrc", "cn_state.h"),
            os.path.join(_HERE, "..", "include", "crowdnav.h")]
    stale = force or not os.path.exists(_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libcn_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(CnConfig)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_blob_bytes.restype = C.c_size_t
        L.orc_blob_bytes.argtypes = [C.c_void_p]
        L.orc_obs_dim.argtypes = [C.c_void_p]
        L.orc_init_blob.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_step.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_clear_done.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_waypoint.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_heading.restype = C.c_float
        L.orc_heading.argtypes = [C.c_void_p] + [C.c_float] * 5
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """E independent worlds stepped on the CPU, same state blob as the GPU library."""

    def __init__(self, cfg: CnConfig, debug: bool = False, threads: int | None = None):
        self.cfg = cfg.copy()
        if threads is not None:
            os.environ["OMP_NUM_THREADS"] = str(threads)
        self._L = lib()
        if threads is not None:
            self._L.orc_set_threads(int(threads))   # (the environment variable only counts before the first parallel region)
        self._ctx = self._L.orc_create(C.byref(self.cfg))
        if not self._ctx:
            raise ValueError("oracle rejected the config")
        self.E = cfg.n_envs
        self.D = cfg.obs_dim
        self.NR = cfg.n_samples - 1
        self.blob = np.zeros(self._L.orc_blob_bytes(self._ctx) // 4, dtype=np.uint32)
        self._L.orc_init_blob(self._ctx, _p(self.blob))
        self.obs = np.zeros((self.E, self.D), dtype=np.float32)
        self.reward = np.zeros(self.E, dtype=np.float32)
        self.done = np.zeros(self.E, dtype=np.uint8)
        self.ranges = np.zeros((self.E, self.NR), dtype=np.float32) if debug else None
        self.hit_ids = np.zeros((self.E, self.NR), dtype=np.uint8) if debug else None

    def __del__(self):
        try:
            if self._ctx:
                self._L.orc_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    def reset(self, mask: np.ndarray | None = None) -> np.ndarray:
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self._L.orc_reset(self._ctx, _p(self.blob), _p(mask), _p(self.obs), _p(self.ranges), _p(self.hit_ids))
        return self.obs

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.E, 2)
        self._L.orc_step(self._ctx, _p(self.blob), _p(a), _p(self.obs), _p(self.reward), _p(self.done),
                         _p(self.ranges), _p(self.hit_ids))
        return self.obs, self.reward, self.done

    def clear_done(self, mask: np.ndarray | None = None) -> None:
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self._L.orc_clear_done(self._ctx, _p(self.blob), _p(mask))

    def counters(self) -> np.ndarray:
        out = np.zeros((self.E, 4), dtype=np.int32)
        self._L.orc_counters(self._ctx, _p(self.blob), _p(out))
        return out

    # -- state views (blob layout: crowdnav_b200/csrc/cn_state.h) -------------
    def robot_words(self) -> np.ndarray:
        return self.blob[16:16 + self.E * 16].reshape(self.E, 16)

    def ped_a(self) -> np.ndarray:
        n = self.cfg.n_peds
        o = 16 + self.E * 16
        return self.blob[o:o + self.E * n * 4].reshape(self.E, n, 4)

    def ped_b(self) -> np.ndarray:
        n = self.cfg.n_peds
        o = 16 + self.E * 16 + self.E * n * 4
        return self.blob[o:o + self.E * n * 4].reshape(self.E, n, 4)

    def trk(self) -> np.ndarray:
        """[E, 396] tracker records (CN_FLAG_RISK_FAITHFUL; crowdnav_b200/csrc/cn_faithful_state.h)."""
        n = self.cfg.n_peds
        o = 16 + self.E * 16 + 2 * self.E * n * 4
        return self.blob[o:o + self.E * 396].reshape(self.E, 396)

    def robot_pose(self) -> np.ndarray:
        """[E, 3] float64 x, y, yaw decoded from the fixed-point state."""
        r = self.robot_words()
        x = r[:, 0].view(np.int32).astype(np.float64) / 2.0 ** 24
        y = r[:, 1].view(np.int32).astype(np.float64) / 2.0 ** 24
        yaw = r[:, 2].view(np.int32).astype(np.float64) * (2.0 * np.pi / 2.0 ** 32)
        return np.stack([x, y, yaw], axis=1)

    def ped_xy(self) -> np.ndarray:
        return self.ped_a()[:, :, :2].view(np.int32).astype(np.float64) / 2.0 ** 24

    # -- primitive taps -------------------------------------------------------
    def waypoint(self, x: float, y: float):
        wx, wy = C.c_float(), C.c_float()
        self._L.orc_waypoint(self._ctx, x, y, C.byref(wx), C.byref(wy))
        return wx.value, wy.value

    def heading(self, x, y, yaw, wx, wy) -> float:
        return self._L.orc_heading(self._ctx, x, y, yaw, wx, wy)
