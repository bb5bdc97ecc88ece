#!/usr/bin/env python
"""Where the event-timed duration of a c2 step goes besides the kernel's own span: event pairs around (a) nothing,
(b) a trivial torch kernel, (c) cn_step on 8 worlds (one CTA), (d) cn_step on 4096 worlds; each after the 256 MiB flush
(so the CPU is ahead of the GPU) and after a spin-kernel spacer.  usage: launch_overhead.py   (GPU box)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crowdnav_b200.config import baseline_config
from crowdnav_b200.vec_env import CrowdNavVecEnv
envs = {E: CrowdNavVecEnv(baseline_config(1, n_envs=E, auto_reset=True), device=0) for E in (8, 512, 4096)}
acts = {E: torch.rand((E, 2), device="cuda") * 0.2 for E in envs}
for E, e in envs.items():
    e.reset()
    for _ in range(20): e.step(acts[E])
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
one = torch.zeros(1, device="cuda")
def bracket(fn, spacer, K=40):
    ts = []
    for i in range(K + 5):
        flush.fill_(float(i)) if spacer == "flush" else torch.cuda._sleep(60000)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); fn(); s1.record()
        if i >= 5: ts.append((s0, s1))
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) * 1e3 for a, b in ts])
    return np.median(t), t.min()
for spacer in ("flush", "sleep"):
    print("spacer:", spacer)
    print("  nothing between the events        median %.2f us  min %.2f" % bracket(lambda: None, spacer))
    print("  trivial kernel (1-element fill)    median %.2f us  min %.2f" % bracket(lambda: one.fill_(1.0), spacer))
    for E, e in envs.items():
        print("  cn_step, %4d worlds (tile %2d)     median %.2f us  min %.2f" % ((E, e.kernel_tile) + bracket(lambda: e.step(acts[E]), spacer)))
