#!/usr/bin/env python
"""Per-warp phase timeline of cn_flat_kernel (debug build with -DCN_TIMELINE, %globaltimer stamps).
usage: timeline_flat.py [c2|c3]   (run on the GPU box; libcrowdnav_timeline.so is built by the caller)"""
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
W = env.kernel_tile
print("kernel", env.kernel_name, "tile", W)
n_cta = (E + W - 1) // W
tl = torch.zeros((n_cta * 16 + 16, 16), dtype=torch.int64, device="cuda")
env.reset()
a = torch.zeros((E, 2), device="cuda"); a[:, 0] = 0.15; a[:, 1] = torch.rand(E, device="cuda") * 2 - 1
for _ in range(30):
    env.step(a)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
L.cn_debug_set_timeline_flat.argtypes = [C.c_void_p]
assert L.cn_debug_set_timeline_flat(C.c_void_p(tl.data_ptr())) == 0
names = {14: "kernel entry", 15: "tma issued (thread 0)", 3: "draws done (ped warps)", 4: "ped A done (ped warps)", 0: "after ped barrier A (ped warps)", 11: "B done, before barrier (ped warps)", 1: "tma landed", 13: "peds integrated (ped warps)", 9: "phase 1+2 done (per warp)",
         2: "after #A", 10: "phase 3 done (per warp)", 5: "after #E", 6: "after #F (5)", 12: "phase 6 done (per warp)",
         7: "after #G", 8: "stores drained (thread 0)"}
order = [14, 15, 1, 4, 0, 3, 11, 13, 9, 2, 10, 5, 6, 12, 7, 8]
for it in range(4):
    tl.zero_()
    (flush.fill_(it) if os.environ.get("CN_NOFLUSH") is None else None); torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); env.step(a); s1.record(); torch.cuda.synchronize()
    t = tl.cpu().numpy().astype(np.int64)
    t = t[t[:, 14] > 0]
    t0 = t[:, 14].min()
    rel = (t - t0) / 1000.0
    span = (t.max() - t0) / 1000.0
    print("iter %d: event %.1f us; kernel span by globaltimer %.1f us; warps stamped %d" % (it, s0.elapsed_time(s1) * 1e3, span, len(t)))
    if it < 3: continue
    # who finishes phase 1 (+ 2) when: pose warp 0 (waypoint chain + reward), pose warp 1 (velocities, wall spans, padding),
    # pedestrian warps; and how long each warp spends waiting at the CTA barriers that follow
    wpc = tl.shape[0] and (len(t) // n_cta)
    role = np.arange(len(t)) % wpc if len(t) == n_cta * wpc else None
    if role is not None:
        for nm, sel in (("pose warp 0", role == 0), ("pose warp 1", role == 1), ("pedestrian warps", role >= 2)):
            for k, kn in ((1, "tma landed"), (9, "phase 1+2 done"), (10, "phase 3 done"), (12, "phase 6 done")):
                m = sel & (t[:, k] > 0)
                if m.any():
                    d = (t[m, k] - t[m, 14]) / 1000.0
                    print("   [%-16s] %-16s since CTA start: min %5.2f med %5.2f p90 %5.2f max %5.2f" % (nm, kn, d.min(), np.median(d), np.percentile(d, 90), d.max()))
    for k in order:
        m = t[:, k] > 0
        if not m.any(): continue
        c = rel[m, k]
        d = (t[m, k] - t[m, 14]) / 1000.0    # since this warp's kernel entry
        print("   %-28s abs: med %6.2f max %6.2f | since CTA start: min %5.2f med %5.2f p90 %5.2f max %5.2f" % (
            names[k], np.median(c), c.max(), d.min(), np.median(d), np.percentile(d, 90), d.max()))
    # per-warp durations between consecutive stamps (only warps that have both), by role
    print("   -- per-warp durations (us): median / p90 / max, by role")
    for nm, sel in (("pose warp 0", role == 0), ("pose warp 1", role == 1), ("pedestrian warps", role >= 2)) if role is not None else ():
        prev = 14
        for k in order[1:]:
            m = sel & (t[:, k] > 0) & (t[:, prev] > 0)
            if not m.any():
                continue
            d = (t[m, k] - t[m, prev]) / 1000.0
            print("   [%-16s] %-36s -> %-36s med %5.2f p90 %5.2f max %5.2f (n=%d)" % (nm, names[prev], names[k], np.median(d), np.percentile(d, 90), d.max(), m.sum()))
            prev = k
