"""Configs of the reference-in-the-loop traces (shared by gen_golden.py and the replay test)."""
from crowdnav_b200.config import baseline_config, make_config


def trace_config(name):
    if name == "c1":       # BASELINE configs[0]: 37 samples, K=3, 5 pedestrians, 3 m room
        cfg = baseline_config(0, auto_reset=False)
        cfg.max_steps = 120
        return cfg, {"/turtlebot3/scan_ranges": 37}, dict(n_steps=600, seed=11, seek_goal=False)
    if name == "train":    # the reference's training world: 360 samples, K=8, 14 pedestrians
        cfg = make_config(n_envs=1, auto_reset=False, layout_jitter=0.05, max_steps=150)
        return cfg, {"/turtlebot3/scan_ranges": 360}, dict(n_steps=500, seed=12, seek_goal=False)
    if name == "goal":     # goal seeking: waypoint (+200) / goal (+200) rewards, success episodes
        cfg = make_config(n_envs=1, auto_reset=False, layout_jitter=0.05, max_steps=400, seed=77)
        return cfg, {"/turtlebot3/scan_ranges": 360}, dict(n_steps=700, seed=13, seek_goal=True)
    if name == "test20":   # BASELINE configs[1]'s world: 5 m test room, 20 pedestrians, start (1, 0), goal (-2, 2), steered at the goal
        cfg = baseline_config(1, n_envs=1, auto_reset=False, seed=2024)
        cfg.max_steps = 300
        return cfg, {"/turtlebot3/scan_ranges": 360, "/turtlebot3/desired_pose/x": -2.0, "/turtlebot3/desired_pose/y": 2.0,
                     "/turtlebot3/starting_pose/x": 1.0, "/turtlebot3/starting_pose/y": 0.0}, \
            dict(n_steps=600, seed=16, seek_goal=True)
    if name == "original":         # environment_stage_1_original.py: 363-wide row, goal-relative, its own reward
        cfg = make_config(n_envs=1, auto_reset=False, layout_jitter=0.05, max_steps=150, seed=91, env_original=True)
        return cfg, {"/turtlebot3/scan_ranges": 360}, dict(n_steps=450, seed=14, seek_goal=False)
    if name == "original_goal":    # ... steered at the goal: success episodes (+200)
        cfg = make_config(n_envs=1, auto_reset=False, layout_jitter=0.05, max_steps=400, seed=92, env_original=True)
        return cfg, {"/turtlebot3/scan_ranges": 360}, dict(n_steps=450, seed=15, seek_goal=True)
    raise KeyError(name)


TRACES = ("c1", "train", "goal", "test20")
TRACES_ORIGINAL = ("original", "original_goal")
