#!/usr/bin/env python
"""Per-warp timeline of cn_env_kernel (debug build with -DCN_TIMELINE, %globaltimer stamps).
usage: timeline.py [c2|c3]   (run on the GPU box)"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crowdnav_b200 import _lib
_lib.SO_PATH = os.path.join(ROOT, "crowdnav_b200", "libcrowdnav_timeline.so")
from crowdnav_b200.config import baseline_config
from crowdnav_b200.vec_env import CrowdNavVecEnv
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = baseline_config({"c2": 1, "c3": 2}[wl])
env = CrowdNavVecEnv(cfg, device=0)
L = _lib.load()
E = cfg.n_envs
nw = (E + 13) // 14 * 16
tl = torch.zeros((nw, 16), dtype=torch.int64, device="cuda")
env.reset()
a = torch.zeros((E, 2), device="cuda"); a[:, 0] = 0.15; a[:, 1] = torch.rand(E, device="cuda") * 2 - 1
for _ in range(30):
    env.step(a)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
L.cn_debug_set_timeline.argtypes = [C.c_void_p]
assert L.cn_debug_set_timeline(C.c_void_p(tl.data_ptr())) == 0
res = []
for it in range(5):
    (flush.fill_(it) if os.environ.get("CN_NOFLUSH") is None else None); torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); env.step(a); s1.record(); torch.cuda.synchronize()
    t = tl.cpu().numpy().astype(np.int64)
    t = t[t[:, 5] > 0] if (t[:, 5] > 0).any() else t
    t0 = t[:, 0].min()
    rel = (t - t0) / 1000.0
    names = ["cta start", "tma loaded", "phaseA/P issued(A warps)", "P done", "A visible", "env done", "after barrier", "exit",
             "P: prefilter done", "L: row init done", "L: raster done", "L: count done", "L: clean+min done", "R: risk done", "", ""]
    print("iter %d: event %.1f us; kernel span by globaltimer %.1f us" % (it, s0.elapsed_time(s1) * 1e3, rel[:, 7].max()))
    for k in range(14):
        c = rel[:, k]; c = c[t[:, k] > 0]
        if len(c): print("   %-26s min %6.2f  med %6.2f  p90 %6.2f  max %6.2f" % (names[k], c.min(), np.median(c), np.percentile(c, 90), c.max()))
    d = rel[:, 5] - rel[:, 4]
    print("   env work (A visible -> env done): med %.2f p90 %.2f max %.2f ; P phase med %.2f ; wait-for-A med %.2f" % (
        np.median(d), np.percentile(d, 90), d.max(), np.median(rel[:, 3] - rel[:, 1]), np.median(rel[:, 4] - rel[:, 3])))
