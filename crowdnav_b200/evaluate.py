"""Scenario / evaluation harness: the reference's test protocol on the batched env.

README.md:50-83 of the reference: load a checkpoint, switch learning off, run the four crowd behaviours
(random / towards / crossing / ahead) with 4, 8, 12 or 20 pedestrians in the 5 m test room -- start (1, 0),
goal (-2, 2), `min_scan_range 0.0` so that a close pass is scored, not terminal -- and log per episode
`episode_number, success_episode, failure_episode, episode_reward, episode_step, ego_safety_score,
social_safety_score, timelapse` (utils.record_data, UTL:53-64; start_td3_training.py:144-161).

Here the episodes of one scenario run in parallel worlds on the GPU; `timelapse` is simulated time
(steps * dt), since there is no wall-clock coupling any more.
"""
from __future__ import annotations

import csv
from typing import Dict, List, Sequence

import torch

from .config import (LAYOUT_TEST_20, TABLE_AHEAD_20, TABLE_CROSSING_20, TABLE_TOWARDS_20, CnConfig, behavior_random,
                     behavior_table, test_world_20)

CSV_HEADER = ["episode_number", "success_episode", "failure_episode", "episode_reward", "episode_step",
              "ego_safety_score", "social_safety_score", "timelapse"]          # UTL:56-57

# which of the 20 world-file pedestrians each crowd size uses (simulate_*_{4,8,12,20}.py, turtlebot3_obstacle_N.world)
SCENARIO_IDS = {4: [4, 5, 9, 11], 8: [1, 3, 4, 5, 7, 9, 11, 12], 12: list(range(1, 13)), 20: list(range(1, 21))}
_TABLES = {"towards": TABLE_TOWARDS_20, "crossing": TABLE_CROSSING_20, "ahead": TABLE_AHEAD_20}


def scenario_config(behavior: str, n_peds: int, fast: bool = False, n_envs: int = 256, k_obstacles: int = 8,
                    max_steps: int = 1000, seed: int = 1234, **kw) -> CnConfig:
    """One cell of the README's evaluation grid as a cn_config.

    speeds: 0.1 m/s (0.2 for the `_fast` scripts, 0.04 for the 20-pedestrian ones); random crowds hold a draw for
    2.25 s (11.25 s with 20 pedestrians)."""
    if n_peds not in SCENARIO_IDS:
        raise ValueError("crowd sizes of the reference: 4, 8, 12, 20")
    ids = SCENARIO_IDS[n_peds]
    layout = [LAYOUT_TEST_20[i - 1] for i in ids]
    speed = 0.04 if n_peds == 20 else (0.2 if fast else 0.1)
    if behavior == "random":
        beh = behavior_random(speed, 11.25 if n_peds == 20 else 2.25)
    elif behavior in _TABLES:
        beh = behavior_table([_TABLES[behavior][i - 1] for i in ids], speed, 0.5)
    else:
        raise ValueError("behaviour must be random, towards, crossing or ahead")
    kw.setdefault("collision_range", 0.0)            # README.md:60-62
    return test_world_20(n_envs=n_envs, n_peds=n_peds, layout=layout, behaviors=[beh], k_obstacles=k_obstacles,
                         max_steps=max_steps, seed=seed, auto_reset=True, **kw)


@torch.no_grad()
def evaluate(env, actor, n_episodes: int, max_launches: int = 100000) -> List[Dict[str, float]]:
    """Run `actor` greedily (no exploration noise: README "learning = False") for n_episodes episodes.

    `env` is a CrowdNavVecEnv built with auto_reset.  Every world gets a fixed QUOTA of ceil(n_episodes / E)
    episodes: exactly its first `quota` episodes are recorded, later ones are ignored, and the run lasts until every
    world has met its quota.  (Keeping "the first n episodes to finish" across auto-resetting parallel worlds would
    over-represent short episodes -- under this protocol short means success -- and cut the time-outs off.)
    Returns one dict per recorded episode with the CSV columns, ordered by world, then by episode; the list has
    quota * E >= n_episodes entries so that no world is favoured."""
    E, dev = env.E, env.device
    quota = max(1, -(-n_episodes // E))
    obs = env.reset().clone()
    ret = torch.zeros(E, device=dev)
    length = torch.zeros(E, device=dev)
    n_rec = torch.zeros(E, dtype=torch.int64, device=dev)           # episodes recorded per world
    # per (world, episode slot): success, ego, social, present, return, length
    table = torch.zeros((E, quota, 6), dtype=torch.float32, device=dev)
    dt = float(env.cfg.dt)
    world_idx = torch.arange(E, device=dev)
    for it in range(max_launches):
        nobs, r, d = env.step(actor(obs).contiguous())
        live = d != 2
        ret += torch.where(live, r, torch.zeros_like(r))
        length += live.float()
        ended = d == 1
        c = env.counters().float()                                  # success, ego, social, obstacle-present
        rec = ended & (n_rec < quota)
        slot = n_rec.clamp(max=quota - 1)
        row = torch.cat([c, ret.unsqueeze(1), length.unsqueeze(1)], 1)
        cur = table[world_idx, slot]
        table[world_idx, slot] = torch.where(rec.unsqueeze(1), row, cur)
        n_rec += rec.to(torch.int64)
        ret.masked_fill_(ended, 0.0)
        length.masked_fill_(ended, 0.0)
        obs = nobs.clone()
        if (it & 15) == 15 and bool((n_rec >= quota).all()):        # one read-back every 16 launches
            break
    t = table.cpu()
    done_n = n_rec.cpu()
    rows: List[Dict[str, float]] = []
    for w in range(E):
        for k in range(int(done_n[w])):
            succ, ego, soc, pres, rr, ll = (float(x) for x in t[w, k])
            # ENV:1269-1283; the reference divides by zero when no obstacle was ever seen -- reported as 1.0 here
            rows.append({"episode_number": len(rows) + 1, "success_episode": bool(succ), "failure_episode": not bool(succ),
                         "episode_reward": rr, "episode_step": int(ll),
                         "ego_safety_score": 1.0 - ego / pres if pres > 0 else 1.0,
                         "social_safety_score": 1.0 - soc / pres if pres > 0 else 1.0,
                         "timelapse": ll * dt})
    return rows


def write_csv(rows: Sequence[Dict[str, float]], path: str) -> None:
    """utils.record_data's file format (UTL:53-64)."""
    with open(path, "w", newline="") as fp:
        w = csv.DictWriter(fp, fieldnames=CSV_HEADER, delimiter=",", lineterminator="\n")
        w.writeheader()
        for r in rows:
            w.writerow(r)


def summarize(rows: Sequence[Dict[str, float]]) -> Dict[str, float]:
    n = max(len(rows), 1)
    return {"episodes": len(rows), "success_rate": sum(r["success_episode"] for r in rows) / n,
            "mean_reward": sum(r["episode_reward"] for r in rows) / n, "mean_steps": sum(r["episode_step"] for r in rows) / n,
            "ego_safety": sum(r["ego_safety_score"] for r in rows) / n,
            "social_safety": sum(r["social_safety_score"] for r in rows) / n}
