"""Scenario / evaluation harness."""
import csv
import os

import pytest

from crowdnav_b200.evaluate import CSV_HEADER, SCENARIO_IDS, scenario_config, summarize, write_csv


def test_scenario_grid_matches_the_reference_scripts():
    """crowd sizes / ids / speeds / goal of README.md:60-83 and simulate_*_{4,8,12,20}.py."""
    for beh in ("random", "towards", "crossing", "ahead"):
        for n in (4, 8, 12, 20):
            cfg = scenario_config(beh, n, n_envs=8)
            assert cfg.n_peds == n and (cfg.goal_x, cfg.goal_y) == (-2.0, 2.0) and cfg.collision_range == 0.0
            assert abs(cfg.behavior_speed[0] - (0.04 if n == 20 else 0.1)) < 1e-7
    cfg = scenario_config("towards", 4, fast=True, n_envs=8)
    assert abs(cfg.behavior_speed[0] - 0.2) < 1e-7
    # towards_4: obstacle_4 (+,+), 5 (+,0), 9 (+,-), 11 (+,-)   (simulate_towards_4.py:89-92)
    t = [(cfg.behavior_table[0][i][0], cfg.behavior_table[0][i][1]) for i in range(4)]
    assert t == [(1.0, 1.0), (1.0, 0.0), (1.0, -1.0), (1.0, -1.0)]
    # turtlebot3_obstacle_4.world: obstacle_4 at (-1.28, -0.75)
    assert abs(cfg.ped_layout[0][0] + 1.28) < 1e-6 and abs(cfg.ped_layout[0][1] + 0.75) < 1e-6
    assert SCENARIO_IDS[8] == [1, 3, 4, 5, 7, 9, 11, 12]


def test_csv_format(tmp_path):
    rows = [{"episode_number": 1, "success_episode": True, "failure_episode": False, "episode_reward": 123.0,
             "episode_step": 77, "ego_safety_score": 1.0, "social_safety_score": 0.9, "timelapse": 11.55}]
    p = os.path.join(tmp_path, "td3_training.csv")
    write_csv(rows, p)
    got = list(csv.reader(open(p)))
    assert got[0] == CSV_HEADER and got[1][0] == "1" and got[1][4] == "77"
    assert summarize(rows)["success_rate"] == 1.0


@pytest.mark.gpu
def test_evaluation_run_on_gpu(tmp_path):
    import torch
    from crowdnav_b200.evaluate import evaluate
    from crowdnav_b200.rollout import load_reference_actor
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    actor = load_reference_actor(os.path.join(os.path.dirname(__file__), "golden", "td3_actor_k8_ep2500.npz"), "cuda")
    env = CrowdNavVecEnv(scenario_config("crossing", 8, n_envs=512, max_steps=300), device=0)
    rows = evaluate(env, actor, 600)
    assert len(rows) == 1024            # quota of ceil(600 / 512) = 2 episodes for EVERY world (no length bias)
    s = summarize(rows)
    assert 0.0 <= s["ego_safety"] <= 1.0 and 0.0 <= s["social_safety"] <= 1.0 and s["mean_steps"] <= 300
    assert all(r["episode_step"] <= 300 for r in rows)
    write_csv(rows, os.path.join(tmp_path, "eval.csv"))
