#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE (see
tests/ref_harness.py for how it is imported without ROS).  Run in the build
container, where /root/reference exists:

    python tests/gen_golden.py

Fixtures (all small, committed):
  utils_vectors.npz   inputs/outputs of the shapely-free functions of utils.py
  env_methods.npz     Env.get_heading_to_goal / get_distance_to_goal / goal boxes /
                      compute_reward on a grid of states (reward truth table)
  waypoints.npz       utils.get_local_goal_waypoints (through the shapely stand-in)
  td3_actor_k8_ep2500.npz
                      the shipped K=8 TD3 actor's state_dict (checkpoint drop-in fixture)
  trace_original.npz, trace_original_goal.npz
                      the same for environment_stage_1_original.Env (CN_FLAG_ENV_ORIGINAL)
  trace_c1.npz, trace_train.npz, trace_goal.npz
                      reference-in-the-loop traces: the reference's Env.reset/step
                      (get_state + compute_reward, unmodified) observing physics
                      injected from the CPU oracle's simulator; per step the
                      action, the injected odometry + raw scan and the
                      reference's (state, reward, done).
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_harness import DEFAULT_PARAMS, Reference, Scan, _XYZ, _Quat  # noqa: E402
from crowdnav_b200.config import baseline_config, make_config  # noqa: E402
from oracle.oracle import OracleEnv  # noqa: E402
from trace_configs import TRACES, TRACES_ORIGINAL, trace_config  # noqa: E402

OUT = os.path.join(HERE, "golden")


def gen_utils(ref):
    u = ref.utils
    rng = np.random.default_rng(0)
    # get_scan_ranges (UTL:375-392)
    raw = rng.uniform(0.05, 0.9, size=(6, 360))
    raw[0, ::7] = np.inf
    raw[1, ::5] = np.nan
    raw[2, ::9] = 0.0
    raw[3, :] = np.inf
    cleaned = np.array([u.get_scan_ranges(Scan(r.tolist()), 360, 0.6) for r in raw])
    # convert_laserscan_to_coordinate (UTL:110-126); Python-2 `360 / 359` == 1 is reproduced
    # by passing max_angle = 359 (so the true quotient is the Py2 integer quotient)
    scans = rng.uniform(0.08, 0.6, size=(5, 359))
    poses = rng.uniform(-2, 2, size=(5, 3))
    coords = np.array([u.convert_laserscan_to_coordinate(s.tolist(), 360, _XYZ(p[0], p[1]), p[2], 359)
                       for s, p in zip(scans, poses)])
    scans37 = rng.uniform(0.08, 0.6, size=(5, 36))
    coords37 = np.array([u.convert_laserscan_to_coordinate(s.tolist(), 37, _XYZ(p[0], p[1]), p[2], 360)
                         for s, p in zip(scans37, poses)])
    ttc = np.concatenate([rng.uniform(-2, 5, 200), [0.15, 1.5, 0.01, -0.1]])
    cp_ttc = np.array([u.compute_collision_prob(float(t)) for t in ttc] + [u.compute_collision_prob(None)])
    d = np.concatenate([rng.uniform(0.0, 0.8, 200), [0.12, 0.36, 0.6, 0.61]])
    cp_dto = np.array([u.compute_general_collision_prob(float(x), 0.6, 0.12) for x in d])
    nscan = np.array([u.estimate_num_obs_scans(float(x), 0.6, 0.12) for x in d])
    ring = np.array(u.convert_laserscan_to_coordinate([0.6] * 359, 360, _XYZ(0.3, -0.2), 1.1, 359))
    bbox = u.compute_average_bounding_box_size(ring.tolist())
    p2 = rng.uniform(-1, 1, size=(50, 2, 2))
    tv = np.array([u.get_timestep_velocity(p.tolist(), 0.15) for p in p2])
    yaw = rng.uniform(-math.pi, math.pi, 50)
    y360 = np.array([u.convert_yaw_to_360deg(float(y)) for y in yaw])
    np.savez_compressed(os.path.join(OUT, "utils_vectors.npz"), raw=raw, cleaned=cleaned, scans=scans, poses=poses,
                        coords=coords, scans37=scans37, coords37=coords37, ttc=ttc, cp_ttc=cp_ttc, d=d, cp_dto=cp_dto,
                        nscan=nscan, bbox=np.array([bbox]), p2=p2, tv=tv, yaw=yaw, y360=y360)
    print("utils_vectors: bbox", bbox)


def gen_env_methods(ref):
    rng = np.random.default_rng(1)
    env = ref.make_env()
    n = 400
    pose = np.column_stack([rng.uniform(-1.4, 1.4, n), rng.uniform(-1.4, 1.4, n), rng.uniform(-math.pi, math.pi, n)])
    wp = np.column_stack([rng.uniform(-1.4, 1.4, n), rng.uniform(-1.4, 1.4, n)])
    head, dist, in_wp, in_goal = [], [], [], []
    for p, w in zip(pose, wp):
        env.waypoint_desired_point.x, env.waypoint_desired_point.y = float(w[0]), float(w[1])
        head.append(env.get_heading_to_goal(_XYZ(p[0], p[1]), _Quat(p[2])))
        dist.append(env.get_distance_to_goal(_XYZ(p[0], p[1])))
        in_wp.append(env.is_in_desired_position(_XYZ(p[0], p[1])))
        in_goal.append(env.is_in_true_desired_position(_XYZ(p[0], p[1])))
    # box edges (half-open, ENV:1296-1297)
    edge_pts = np.array([[-0.8, 1.0], [-1.2, 1.0], [-1.0, 1.2], [-1.0, 0.8], [-0.8000001, 1.1999999], [-1.1999999, 0.8000001]])
    edge_goal = [env.is_in_true_desired_position(_XYZ(p[0], p[1])) for p in edge_pts]

    # compute_reward truth table (ENV:1046-1162): heading/distance before and after, done / not done,
    # robot outside any waypoint box so the shapely-dependent refresh is not entered here
    vals = [-1.5, -0.2, 0.0, 0.2, 1.5]
    rows = []
    for ph in vals:
        for ch in vals:
            for pd, cd in ((1.0, 0.9), (1.0, 1.0), (1.0, 1.1)):
                for done, at_goal in ((False, False), (True, False), (True, True)):
                    e = ref.make_env()
                    e.previous_heading, e.previous_distance = ph, pd
                    e.waypoint_desired_point.x, e.waypoint_desired_point.y = 5.0, 5.0
                    pos = (-1.0, 1.0) if at_goal else (0.3, 0.3)
                    ref.set_odom(e, pos[0], pos[1], 0.0, 0.0, 0.0)
                    state = [0.6] * 359 + [ch, cd] + [0.0] * 37
                    r, d = e.compute_reward(state, 5, done)
                    rows.append([ph, ch, pd, cd, float(done), float(at_goal), float(r), float(d),
                                 float(e.episode_success), float(e.episode_failure)])
    np.savez_compressed(os.path.join(OUT, "env_methods.npz"), pose=pose, wp=wp, head=np.array(head), dist=np.array(dist),
                        in_wp=np.array(in_wp), in_goal=np.array(in_goal), edge_pts=edge_pts, edge_goal=np.array(edge_goal),
                        reward_table=np.array(rows))
    print("env_methods: reward rows", len(rows))


def gen_waypoints(ref):
    rng = np.random.default_rng(2)
    n = 500
    agent = np.column_stack([rng.uniform(-1.4, 1.4, n), rng.uniform(-1.4, 1.4, n)])
    agent[:20] = np.array([-1.0, 1.0]) + rng.uniform(-0.29, 0.29, size=(20, 2)) * 0.7     # inside the ring: fallback
    goal = (-1.0, 1.0)
    out = np.array([ref.utils.get_local_goal_waypoints([float(a[0]), float(a[1])], list(goal), 0.3) for a in agent])
    np.savez_compressed(os.path.join(OUT, "waypoints.npz"), agent=agent, goal=np.array(goal), waypoint=out)
    d = np.hypot(out[:, 0] - agent[:, 0], out[:, 1] - agent[:, 1])
    print("waypoints: ring distance range", d[20:].min(), d[20:].max())


def raw_scan_from_oracle(o: OracleEnv) -> np.ndarray:
    """Rebuild the Gazebo-order LaserScan.ranges (R samples, +inf = no return) from the
    oracle's debug taps (obs order j = R-1-i, sample 0 dropped)."""
    R = o.cfg.n_samples
    raw = np.full(R, np.inf, dtype=np.float64)
    rng_obs, hid = o.ranges[0], o.hit_ids[0]
    for i in range(1, R):
        j = (R - 1) - i
        if hid[j] != 0xFF:
            raw[i] = float(rng_obs[j])
    return raw


def oracle_odom(o: OracleEnv):
    r = o.robot_words()[0]
    x = float(np.float32(np.int32(r[0])) * np.float32(2.0 ** -24))
    y = float(np.float32(np.int32(r[1])) * np.float32(2.0 ** -24))
    yaw = float(np.float32(np.int32(r[2])) * np.float32(1.4629180792671596e-09))
    v = float(r[3:4].view(np.float32)[0])
    w = float(r[4:5].view(np.float32)[0])
    return x, y, yaw, v, w


def gen_trace(ref, name, cfg, n_steps, seed, params, seek_goal=False):
    """Reference Env in the loop, physics from the oracle's simulator."""
    ref.params.clear()
    ref.params.update(DEFAULT_PARAMS)       # every trace starts from the YAML defaults
    ref.params.update(params)
    o = OracleEnv(cfg, debug=True)
    rng = np.random.default_rng(seed)
    R, K = cfg.n_samples, cfg.k_obstacles
    rec = {k: [] for k in ("action", "odom", "scan", "ref_state", "ref_reward", "ref_done", "episode_start",
                           "ref_success", "oracle_obs")}
    state_holder = {}

    original = bool(cfg.flags & 4)         # CN_FLAG_ENV_ORIGINAL: the reference's environment_stage_1_original.Env

    def new_episode():
        o.reset()
        env = (ref.make_env_original(max_step=cfg.max_steps) if original
               else ref.make_env(max_step=cfg.max_steps, k_obstacle_count=K))
        ref.set_odom(env, *oracle_odom(o))
        state_holder["scan"] = raw_scan_from_oracle(o)
        s = env.reset()
        env.done = False
        return env, np.asarray(s, dtype=np.float64)

    ref.scan_source = lambda: Scan(state_holder["scan"].tolist())
    env, s0 = new_episode()
    rec["action"].append([0.0, 0.0]); rec["odom"].append(oracle_odom(o)); rec["scan"].append(state_holder["scan"].copy())
    rec["ref_state"].append(s0); rec["ref_reward"].append(0.0); rec["ref_done"].append(0.0); rec["episode_start"].append(1.0)
    rec["ref_success"].append(0.0); rec["oracle_obs"].append(o.obs[0].copy())
    step = 0
    for t in range(n_steps):
        a = np.array([[rng.uniform(0.0, 0.22), rng.uniform(-2.0, 2.0)]], dtype=np.float32)
        if seek_goal:                              # steer at the goal: waypoint / goal rewards, success episodes
            x, y, yaw, _, _ = oracle_odom(o)
            err = math.atan2(cfg.goal_y - y, cfg.goal_x - x) - yaw
            err = (err + math.pi) % (2.0 * math.pi) - math.pi
            a[0, 0] = 0.22 if abs(err) < 0.8 else 0.05
            a[0, 1] = np.float32(max(-2.0, min(2.0, 2.5 * err)) + rng.uniform(-0.2, 0.2))
        elif rng.uniform() < 0.5:                  # bias toward driving forward so episodes end in all three ways
            a[0, 0] = 0.22
            a[0, 1] = np.float32(rng.uniform(-0.6, 0.6))

        def physics(dt):
            o.step(a)
            ref.set_odom(env, *oracle_odom(o))
            state_holder["scan"] = raw_scan_from_oracle(o)
        ref.clock.on_sleep = physics
        s, r, d = env.step([float(a[0, 0]), float(a[0, 1])], step + 1, mode="continuous")
        ref.clock.on_sleep = None
        step += 1
        rec["action"].append(a[0].astype(np.float64)); rec["odom"].append(oracle_odom(o))
        rec["scan"].append(state_holder["scan"].copy()); rec["ref_state"].append(np.asarray(s, dtype=np.float64))
        rec["ref_reward"].append(float(r)); rec["ref_done"].append(float(d)); rec["episode_start"].append(0.0)
        rec["ref_success"].append(float(env.episode_success)); rec["oracle_obs"].append(o.obs[0].copy())
        if d:
            env, s0 = new_episode()
            step = 0
            rec["action"].append([0.0, 0.0]); rec["odom"].append(oracle_odom(o)); rec["scan"].append(state_holder["scan"].copy())
            rec["ref_state"].append(s0); rec["ref_reward"].append(0.0); rec["ref_done"].append(0.0)
            rec["episode_start"].append(1.0); rec["ref_success"].append(0.0); rec["oracle_obs"].append(o.obs[0].copy())
    arrays = {k: np.asarray(v) for k, v in rec.items()}
    arrays["scan"] = arrays["scan"].astype(np.float32)
    arrays["ref_state"] = arrays["ref_state"].astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "trace_%s.npz" % name), **arrays)
    n_ep = int(arrays["episode_start"].sum())
    succ = int(((arrays["ref_done"] > 0) & (arrays["ref_success"] > 0)).sum())
    print("trace_%s: %d rows, %d episodes, %d successful, reward range [%g, %g]" % (
        name, len(arrays["action"]), n_ep, succ, arrays["ref_reward"].min(), arrays["ref_reward"].max()))


def gen_actor_fixture():
    """The shipped K=8 TD3 actor (398 -> 256 -> 256 -> 2) as an npz of its state_dict tensors: the checkpoint
    drop-in fixture of SURVEY.md section 4 (weights are data the reference ships, not code)."""
    import torch
    src = "/root/reference/turtlebot3_rl_sim/src/models/td3/turtlebot3_top_8_obstacle/td3_actor_model_ep2500.pt"
    sd = torch.load(src, map_location="cpu")
    np.savez_compressed(os.path.join(OUT, "td3_actor_k8_ep2500.npz"), **{k: v.numpy() for k, v in sd.items()})
    print("td3_actor_k8_ep2500: ", {k: tuple(v.shape) for k, v in sd.items()})


def main():
    """All fixtures, or only the traces named on the command line (`python tests/gen_golden.py original ...`)."""
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    ref = Reference()
    if not only:
        gen_actor_fixture()
        gen_utils(ref)
        gen_env_methods(ref)
        gen_waypoints(ref)
    for name in TRACES + TRACES_ORIGINAL:
        if only and name not in only:
            continue
        cfg, params, kw = trace_config(name)
        gen_trace(ref, name, cfg, kw["n_steps"], kw["seed"], params, seek_goal=kw["seek_goal"])

if __name__ == "__main__":
    main()
