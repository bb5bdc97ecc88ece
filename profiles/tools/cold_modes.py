#!/usr/bin/env python
"""How the step kernel's event-timed duration depends on the way L2 is made cold between steps.
usage: cold_modes.py [c2|c3]   (GPU box).  Modes:
  flush   : 256 MiB fill before every step (leaves L2 full of dirty lines, evicts the kernel's code too)
  rotate  : R independent replicas of the batch (> 2x L2 in total) stepped round-robin, a spin kernel as spacer
  warm    : one batch stepped back to back (state L2-resident), a spin kernel as spacer
The spacer (torch.cuda._sleep) keeps the CPU ahead of the GPU so the event pair brackets GPU time only."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crowdnav_b200.config import baseline_config
from crowdnav_b200.vec_env import CrowdNavVecEnv
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = baseline_config({"c2": 1, "c3": 2, "c5": 4}[wl], auto_reset=True)
E = cfg.n_envs
per = E * (64 + 32 * cfg.n_peds + 4 * ((cfg.n_samples - 1) + 7 + 4 * cfg.k_obstacles))
R = max(2, int(np.ceil(300e6 / per)))
envs = [CrowdNavVecEnv(cfg, device=0) for _ in range(R)]
print("kernel", envs[0].kernel_name, "tile", envs[0].kernel_tile, "replicas", R, "MB each %.1f" % (per / 1e6))
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [torch.stack([torch.rand(E, device="cuda", generator=g) * 0.22, torch.rand(E, device="cuda", generator=g) * 4 - 2], 1).contiguous() for _ in range(16)]
for e in envs:
    e.reset()
    for i in range(20): e.step(acts[i % 16])
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
def run(mode, K=60):
    evs = []
    for i in range(K + 5):
        env = envs[i % R] if mode == "rotate" else envs[0]
        if mode == "flush": flush.fill_(float(i))
        else: torch.cuda._sleep(60000)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); env.step(acts[i % 16]); s1.record()
        if i >= 5: evs.append((s0, s1))
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) * 1e3 for a, b in evs])
    print("%-7s mean %.2f us  median %.2f  min %.2f  max %.2f" % (mode, t.mean(), np.median(t), t.min(), t.max()))
for m in ("flush", "rotate", "warm", "flush", "rotate", "warm"):
    run(m)
