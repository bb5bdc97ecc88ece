#!/usr/bin/env python
"""A few steps of small batches through every kernel, for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python profiles/tools/sanitize_small.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crowdnav_b200.config import baseline_config
from crowdnav_b200.vec_env import CrowdNavVecEnv

which = sys.argv[1] if len(sys.argv) > 1 else "flat"
rng = np.random.default_rng(0)
cases = {"flat": [(1, 40, 0), (4, 12, 0), (0, 3, 0)], "faithful": [(1, 12, 8)], "warp": [(1, 28, 0)]}[which]
if which == "warp":
    os.environ["CN_KERNEL"] = "warp"
for idx, E, flags in cases:
    cfg = baseline_config(idx, n_envs=E, auto_reset=True)
    cfg.flags |= flags
    env = CrowdNavVecEnv(cfg, device=0)
    env.reset()
    for t in range(6):
        a = np.stack([rng.uniform(0.1, 0.22, E), rng.uniform(-1, 1, E)], 1).astype(np.float32)
        env.step(torch.from_numpy(a).cuda())
    m = torch.zeros(E, dtype=torch.uint8, device="cuda"); m[::2] = 1
    env.reset(m)
    env.counters()
    torch.cuda.synchronize()
    print(which, "config", idx, "worlds", E, "kernel", env.kernel_name, "tile", env.kernel_tile, "ok", flush=True)
    env.close()
