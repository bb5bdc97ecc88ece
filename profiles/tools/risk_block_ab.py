#!/usr/bin/env python
"""Does the policy care which risk block feeds it?  The reference's shipped K=8 TD3 actor (tests/golden fixture) driven
greedily in the same worlds with `risk_intended` (ideal association) and `risk_faithful` (the reference's own
segmentation / tracker).  Runs on the CPU oracle (test infrastructure), so it needs no GPU.
usage: python profiles/tools/risk_block_ab.py [n_worlds]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crowdnav_b200.config import CN_FLAG_RISK_FAITHFUL, shipped_actor_world
from crowdnav_b200.rollout import load_reference_actor
from oracle.oracle import OracleEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
actor = load_reference_actor(os.path.join(ROOT, "tests", "golden", "td3_actor_k8_ep2500.npz"))
for name, flag in (("risk_intended", 0), ("risk_faithful", CN_FLAG_RISK_FAITHFUL)):
    cfg = shipped_actor_world(n_envs=E, max_steps=600)
    cfg.flags |= flag
    env = OracleEnv(cfg)
    NR, K = cfg.n_samples - 1, cfg.k_obstacles
    obs = env.reset().copy()
    alive = np.ones(E, bool)
    succ = steps = occ = rows = 0
    ret = np.zeros(E)
    for t in range(600):
        with torch.no_grad():
            a = actor(torch.from_numpy(obs)).numpy().astype(np.float32)
        o, r, d = env.step(a)
        blk = o[alive][:, NR + 7:].reshape(-1, K, 4)
        xy = o[alive][:, NR + 2:NR + 4]
        occ += int((np.abs(blk[:, :, :2] - xy[:, None, :]).max(2) > 0).sum()); rows += int(alive.sum())
        ret[alive] += r[alive]
        ended = alive & (d > 0)
        succ += int(env.counters()[ended, 0].sum()); steps += int(alive.sum())
        alive &= ~ended
        obs = o.copy()
        if not alive.any():
            break
    c = env.counters()
    pres = np.maximum(c[:, 3], 1)
    print("%-14s worlds %d  success %.3f  mean return %.1f  mean steps %.1f  occupied K slots per row %.3f  "
          "ego safety %.3f  social safety %.3f" % (name, E, succ / E, ret.mean(), steps / E, occ / max(rows, 1),
                                                    (1 - c[:, 1] / pres).mean(), (1 - c[:, 2] / pres).mean()))
