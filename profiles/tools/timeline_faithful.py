#!/usr/bin/env python
"""Per-world stage timeline of cn_faithful_kernel (debug build with -DCN_TIMELINE, %globaltimer stamps: stage k is
stamped by the world's thread 0 after the barrier that ends it).
Build first:  nvcc <flags of crowdnav_b200/_lib.py> -DCN_TIMELINE -o crowdnav_b200/libcrowdnav_timeline.so crowdnav_b200/csrc/*.cu
usage: timeline_faithful.py [c2|c3]   (GPU box)"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crowdnav_b200 import _lib
_lib.SO_PATH = os.path.join(ROOT, "crowdnav_b200", "libcrowdnav_timeline.so")
from crowdnav_b200.config import CN_FLAG_RISK_FAITHFUL, baseline_config
from crowdnav_b200.vec_env import CrowdNavVecEnv
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = baseline_config({"c2": 1, "c3": 2}[wl])
cfg.flags |= CN_FLAG_RISK_FAITHFUL
env = CrowdNavVecEnv(cfg, device=0)
L = _lib.load()
E = cfg.n_envs
tl = torch.zeros((E + 8, 16), dtype=torch.int64, device="cuda")
env.reset()
a = torch.zeros((E, 2), device="cuda"); a[:, 0] = 0.2; a[:, 1] = torch.rand(E, device="cuda") - 0.5
for _ in range(30):
    env.step(a)
L.cn_debug_set_timeline_faithful.argtypes = [C.c_void_p]
assert L.cn_debug_set_timeline_faithful(C.c_void_p(tl.data_ptr())) == 0
names = ["entry", "hit points", "gradients", "candidate compaction", "typing walk (thread 0)", "association + hit flags",
         "segment ends combined", "sub-segment ends counted", "offsets + running counts", "confirmation",
         "verdicts compacted (thread 0)", "best IoU per tracked", "tracker update (thread 0)", "collision cone",
         "CP / ranking / row (thread 0)"]
for it in range(3):
    tl.zero_(); torch.cuda.synchronize()
    env.step(a); torch.cuda.synchronize()
    t = tl.cpu().numpy()[:E].astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    d = np.diff(t[:, :15], axis=1) / 1000.0
    life = (t[:, 14] - t[:, 0]) / 1000.0
    print("launch %d: worlds %d, first entry -> last exit %.1f us, world lifetime median %.1f us (p95 %.1f)" % (
        it, len(t), (t[:, 14].max() - t0) / 1000.0, np.median(life), np.percentile(life, 95)))
    for k in range(14):
        print("   %-34s median %6.2f us   p95 %6.2f   share %4.1f %%" % (names[k + 1], np.median(d[:, k]), np.percentile(d[:, k], 95),
                                                                       100.0 * d[:, k].sum() / life.sum()))
