#!/usr/bin/env python
"""The reference's shipped TD3 actors (models/td3/turtlebot3_top_{1,4,8,12,16}_obstacle/td3_actor_model_ep2500.pt),
unchanged, driven greedily in this simulator -- at the reference's OWN settings (training world of
configs/turtlebot3_world.yaml:10-13: goal (-1, 1), dt 0.15 s, instantaneous wheels) and in `shipped_actor_world`
(goal (-0.7, 0.7), dt 0.19 s, wheel ramp), with either risk block.  One episode per world (no auto-reset, so no
length bias); per-episode rows go to an 8-column CSV in the reference's format (UTL:53-64), the summaries to stdout.

Runs on the CPU oracle (test infrastructure; the GPU path is bit-identical to it) and reads the checkpoints from
/root/reference, so it is a GENERATING script for profiles/r02/dropin/, not part of the product or of the tests.
usage: python profiles/tools/dropin_grid.py [n_worlds] [out_dir]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crowdnav_b200.config import CN_FLAG_RISK_FAITHFUL, make_config, shipped_actor_world
from crowdnav_b200.evaluate import summarize, write_csv
from crowdnav_b200.rollout import load_reference_actor
from oracle.oracle import OracleEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02", "dropin")
REF = "/root/reference/turtlebot3_rl_sim/src/models/td3/turtlebot3_top_%d_obstacle/td3_actor_model_ep2500.pt"
MAX_STEPS = 1000                                      # td3.yaml:7
os.makedirs(OUT, exist_ok=True)
lines = []
for K in (1, 4, 8, 12, 16):
    actor = load_reference_actor(REF % K)
    for wname, mk in (("reference_yaml", lambda **kw: make_config(layout_jitter=0.05, **kw)), ("shipped_actor_world", shipped_actor_world)):
        for rname, flag in (("risk_intended", 0), ("risk_faithful", CN_FLAG_RISK_FAITHFUL)):
            cfg = mk(n_envs=E, k_obstacles=K, max_steps=MAX_STEPS, auto_reset=False)
            cfg.flags |= flag
            env = OracleEnv(cfg, threads=os.cpu_count() or 1)
            obs = env.reset().copy()
            alive = np.ones(E, bool)
            ret, length = np.zeros(E), np.zeros(E, int)
            for t in range(MAX_STEPS):
                with torch.no_grad():
                    a = actor(torch.from_numpy(obs)).numpy().astype(np.float32)
                o, r, d = env.step(a)
                ret[alive] += r[alive]
                length[alive] += 1
                alive &= ~(d > 0)
                obs = o.copy()
                if not alive.any():
                    break
            c = env.counters()
            rows = []
            for w in range(E):
                succ, ego, soc, pres = (float(x) for x in c[w])
                rows.append({"episode_number": w + 1, "success_episode": bool(succ), "failure_episode": not bool(succ),
                             "episode_reward": float(ret[w]), "episode_step": int(length[w]),
                             "ego_safety_score": 1.0 - ego / pres if pres > 0 else 1.0,
                             "social_safety_score": 1.0 - soc / pres if pres > 0 else 1.0,
                             "timelapse": float(length[w]) * float(cfg.dt)})
            write_csv(rows, os.path.join(OUT, "td3_top%d_%s_%s.csv" % (K, wname, rname)))
            s = summarize(rows)
            timeouts = int(((length >= MAX_STEPS) & ~np.array([r_["success_episode"] for r_ in rows])).sum())
            line = ("K=%-2d %-20s %-13s episodes %d  success %.3f  time-outs %.3f  mean reward %8.1f  mean steps %6.1f  "
                    "ego safety %.3f  social safety %.3f" % (K, wname, rname, E, s["success_rate"], timeouts / E, s["mean_reward"],
                                                             s["mean_steps"], s["ego_safety"], s["social_safety"]))
            print(line, flush=True)
            lines.append(line)
with open(os.path.join(OUT, "summary.txt"), "w") as f:
    f.write("# shipped TD3 actors, greedy, one episode per world, %d worlds per cell, max %d steps; CPU oracle (bit-identical to the GPU path)\n"
            "# reference_yaml = the training world exactly as committed (goal (-1, 1), dt 0.15 s, instantaneous wheels, 3 m room, 14 pedestrians)\n"
            "# shipped_actor_world = goal (-0.7, 0.7), dt 0.19 s, wheel ramp 1 m/s^2 in 15 sub-steps (crowdnav_b200/config.py)\n" % (E, MAX_STEPS))
    f.write("\n".join(lines) + "\n")
