"""Loader (and in-tree builder) of libcrowdnav.so, the CUDA library behind the
C ABI of include/crowdnav.h.  There is NO CPU fallback: if the library is
missing or a CUDA call fails, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from .config import CnConfig

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
# CN_LIB: an alternative build of the library next to the default one (profiling / A-B experiments only)
SO_PATH = os.path.join(_PKG, os.environ.get("CN_LIB", "libcrowdnav.so"))
SOURCES = ["cn_abi.cu", "cn_step.cu", "cn_flat.cu", "cn_faithful.cu"]
HEADERS = ["cn_math.h", "cn_math64.h", "cn_faithful.h", "cn_faithful_state.h", "cn_state.h", "cn_kernel.h", "cn_dev.h", os.path.join("..", "..", "include", "crowdnav.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # Blackwell B200 only
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                                  # no FMA contraction: cn_math.h is the numeric spec
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",
]


class CrowdNavError(RuntimeError):
    pass


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a without a GPU (about 15 s)."""
    srcs = [os.path.join(_CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(_CSRC, h) for h in HEADERS]
    stale = force or not os.path.exists(SO_PATH) or any(
        os.path.getmtime(d) > os.path.getmtime(SO_PATH) for d in deps if os.path.exists(d))
    if stale:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if not os.path.exists(nvcc):
            nvcc = "nvcc"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH] + srcs
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise CrowdNavError("nvcc failed:\n" + res.stdout)
        if verbose:
            print(res.stdout)
    return SO_PATH


_lib = None


def load() -> C.CDLL:
    """dlopen libcrowdnav.so and declare the ABI.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise CrowdNavError(
            "libcrowdnav.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or crowdnav_b200._lib.build_library(); there is no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    L.cn_last_error.restype = C.c_char_p
    L.cn_last_error.argtypes = []
    L.cn_abi_version.restype = i32
    L.cn_config_default.argtypes = [C.POINTER(CnConfig)]
    L.cn_obs_dim.argtypes = [C.POINTER(CnConfig)]
    L.cn_blob_bytes.restype = sz
    L.cn_blob_bytes.argtypes = [C.POINTER(CnConfig)]
    L.cn_create.argtypes = [C.POINTER(CnConfig), i32, C.POINTER(vp)]
    L.cn_destroy.argtypes = [vp]
    L.cn_reset.argtypes = [vp, vp, vp, vp]
    L.cn_step.argtypes = [vp, vp, vp, vp, vp, vp]
    L.cn_step_n.argtypes = [vp, i32, vp, sz, vp, vp, vp, sz, vp]
    L.cn_graph_create.argtypes = [vp, i32, vp, sz, vp, vp, vp, sz, C.POINTER(vp)]
    L.cn_graph_launch.argtypes = [vp, vp]
    L.cn_graph_destroy.argtypes = [vp]
    L.cn_step_gather.argtypes = [vp, vp, vp, C.POINTER(vp), i32, vp, vp, vp]
    u64 = C.c_uint64
    L.cn_step_gather_signal.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(vp), i32, vp, vp, vp, i32, i32, i32, vp, vp, vp]
    L.cn_step_gather_async.argtypes = [vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), i32, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.cn_gather_flush.argtypes = [vp, i32, vp, C.POINTER(vp), C.POINTER(vp), i32, vp, i32, i32, i32, vp]
    L.cn_gather_decode16.argtypes = [vp, vp, vp, C.c_longlong, C.c_longlong, C.c_longlong, vp]
    L.cn_gather_wait.argtypes = [vp, vp, i32, i32, vp]
    L.cn_gather_timeouts.argtypes = [vp, C.POINTER(C.c_uint32), vp]
    L.cn_kernel_ctas.argtypes = [vp]
    L.cn_get_counters.argtypes = [vp, vp, vp]
    L.cn_clear_done.argtypes = [vp, vp, vp]
    L.cn_get_blob.argtypes = [vp, vp, sz, vp]
    L.cn_set_blob.argtypes = [vp, vp, sz, vp]
    L.cn_set_debug_taps.argtypes = [vp, vp, vp]
    L.cn_launch_count.restype = C.c_int64
    L.cn_launch_count.argtypes = [vp]
    L.cn_kernel_name.restype = C.c_char_p
    L.cn_kernel_name.argtypes = [vp]
    L.cn_kernel_tile.argtypes = [vp]
    L.cn_plan_tile.argtypes = [C.POINTER(CnConfig), i32, sz, C.POINTER(i32), C.POINTER(i32), C.POINTER(sz)]
    L.cn_plan_tile_direct.argtypes = [C.POINTER(CnConfig), i32, sz, C.POINTER(i32), C.POINTER(i32), C.POINTER(sz)]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cn_last_error()
        raise CrowdNavError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


ABI_SYMBOLS = [
    "cn_config_default", "cn_obs_dim", "cn_blob_bytes", "cn_create", "cn_destroy", "cn_reset", "cn_step", "cn_step_n", "cn_graph_create", "cn_graph_launch", "cn_graph_destroy", "cn_step_gather", "cn_step_gather_signal", "cn_step_gather_async", "cn_gather_flush", "cn_gather_decode16", "cn_gather_wait", "cn_gather_timeouts", "cn_kernel_ctas",
    "cn_get_counters", "cn_clear_done", "cn_get_blob", "cn_set_blob", "cn_set_debug_taps", "cn_launch_count", "cn_kernel_name", "cn_kernel_tile", "cn_plan_tile", "cn_plan_tile_direct",
    "cn_last_error", "cn_abi_version",
]
