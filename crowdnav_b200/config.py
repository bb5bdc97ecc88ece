"""World / sensor / episode configuration: the ctypes mirror of ``cn_config``
(include/crowdnav.h) plus the reference's worlds as data.

Everything here replaces the reference's rosparam YAML + Gazebo world / xacro
files.  Citations are paths under /root/reference:

* CFG    turtlebot3_rl_sim/src/configs/turtlebot3_world.yaml
* WORLD  turtlebot3_simulations/turtlebot3_gazebo/worlds/turtlebot3_crowd_dense.world
* WORLD20 .../worlds/test_environment/turtlebot3_obstacle_20.world
* XACRO  turtlebot3_description/urdf/turtlebot3_burger.gazebo.xacro
* CROWD  turtlebot3_rl_sim/src/crowd_behaviors/simulate_*.py
"""
from __future__ import annotations

import ctypes as C
import os
import math
from typing import Iterable, Sequence

import numpy as np

CN_MAX_PEDS = 64
CN_MAX_BEHAVIORS = 8
CN_FLAG_AUTO_RESET = 1
CN_FLAG_TOPK_HIGHEST = 2
CN_FLAG_ENV_ORIGINAL = 4     # environment_stage_1_original.py: 363-wide row, goal-relative, its own reward
CN_FLAG_GATHER_WIRE16 = 32   # ... with the 16-bit wire format (tile sized for bulk copies of int16 rows)
CN_FLAG_GATHER_STAGE = 16    # staging tile for the pipelined fused all-gather (cn_step_gather_async)
CN_FLAG_RISK_FAITHFUL = 8    # K block + safety counters from the reference's own segmentation / tracker (ENV:270-1005), float64
CN_BEHAVIOR_RANDOM = 0
CN_BEHAVIOR_TABLE = 1
TICKS_PER_STEP = 3


class CnConfig(C.Structure):
    """Field-for-field mirror of ``struct cn_config``."""

    _fields_ = [
        ("struct_size", C.c_uint32),
        ("flags", C.c_uint32),
        ("n_envs", C.c_int32),
        ("n_peds", C.c_int32),
        ("n_samples", C.c_int32),
        ("k_obstacles", C.c_int32),
        ("max_steps", C.c_int32),
        ("env_id_offset", C.c_int32),
        ("seed", C.c_uint64),
        ("dt", C.c_float),
        ("room_xmin", C.c_float),
        ("room_xmax", C.c_float),
        ("room_ymin", C.c_float),
        ("room_ymax", C.c_float),
        ("start_x", C.c_float),
        ("start_y", C.c_float),
        ("start_yaw", C.c_float),
        ("goal_x", C.c_float),
        ("goal_y", C.c_float),
        ("heading_off_x", C.c_float),
        ("heading_off_y", C.c_float),
        ("max_range", C.c_float),
        ("collision_range", C.c_float),
        ("sensor_min_range", C.c_float),
        ("sensor_sweep", C.c_float),
        ("mount_x", C.c_float),
        ("hit_angle_inc_deg", C.c_float),
        ("ped_radius", C.c_float),
        ("robot_radius", C.c_float),
        ("cp_radius", C.c_float),
        ("waypoint_radius", C.c_float),
        ("goal_box", C.c_float),
        ("rep_strength", C.c_float),
        ("rep_range", C.c_float),
        ("rep_cutoff", C.c_float),
        ("layout_jitter", C.c_float),
        ("wheel_accel", C.c_float),
        ("n_substeps", C.c_int32),
        ("n_behaviors", C.c_int32),
        ("behavior_kind", C.c_int32 * CN_MAX_BEHAVIORS),
        ("behavior_speed", C.c_float * CN_MAX_BEHAVIORS),
        ("behavior_period_ticks", C.c_int32 * CN_MAX_BEHAVIORS),
        ("behavior_stagger_ticks", C.c_int32 * CN_MAX_BEHAVIORS),
        ("behavior_table", (C.c_float * 2) * CN_MAX_PEDS * CN_MAX_BEHAVIORS),
        ("ped_layout", (C.c_float * 2) * CN_MAX_PEDS),
    ]

    # -- convenience ---------------------------------------------------------
    @property
    def obs_dim(self) -> int:
        """(R-1) + 7 + 4K, start_td3_training.py:88; (R-1) + 4 for the original environment's row."""
        if self.flags & CN_FLAG_ENV_ORIGINAL:
            return (self.n_samples - 1) + 4
        return (self.n_samples - 1) + 7 + 4 * self.k_obstacles

    @property
    def n_rays(self) -> int:
        return self.n_samples - 1

    def copy(self) -> "CnConfig":
        out = CnConfig()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(CnConfig))
        return out

    def set_layout(self, poses: Sequence[Sequence[float]]) -> None:
        if len(poses) > CN_MAX_PEDS:
            raise ValueError("at most %d pedestrians" % CN_MAX_PEDS)
        for i in range(CN_MAX_PEDS):
            x, y = poses[i] if i < len(poses) else (0.0, 0.0)
            self.ped_layout[i][0] = x
            self.ped_layout[i][1] = y

    def set_behaviors(self, behaviors: Sequence["Behavior"]) -> None:
        if not 1 <= len(behaviors) <= CN_MAX_BEHAVIORS:
            raise ValueError("1..%d behaviours" % CN_MAX_BEHAVIORS)
        self.n_behaviors = len(behaviors)
        for b, beh in enumerate(behaviors):
            self.behavior_kind[b] = beh.kind
            self.behavior_speed[b] = beh.speed
            self.behavior_period_ticks[b], self.behavior_stagger_ticks[b] = beh.ticks(self.dt if self.dt > 0 else 0.15)
            for n in range(CN_MAX_PEDS):
                dx, dy = beh.table[n] if n < len(beh.table) else (0.0, 0.0)
                self.behavior_table[b][n][0] = dx
                self.behavior_table[b][n][1] = dy


class Behavior:
    """One crowd-mover script as data (SURVEY.md table P')."""

    def __init__(self, kind: int, speed: float, period_s: float, stagger_s: float = 0.1,
                 table: Sequence[Sequence[float]] = ()):
        self.kind = kind
        self.speed = float(speed)
        self.period_s, self.stagger_s = float(period_s), float(stagger_s)
        self.table = [tuple(t) for t in table]

    def ticks(self, dt: float):
        """(period, stagger) in ticks of dt / 3 (0.05 s at the reference's 0.15 s control period)."""
        tick = dt / TICKS_PER_STEP
        return max(1, int(round(self.period_s / tick))), max(0, int(round(self.stagger_s / tick)))


# ---------------------------------------------------------------------------
# World data
# ---------------------------------------------------------------------------
# inner wall faces after the model offset (WORLD:927) and 0.1 m thickness (WORLD:932)
ROOM_3M = (-1.411074, 1.398926, -1.394624, 1.405376)
ROOM_5M = (-2.411074, 2.398926, -2.394624, 2.405376)   # WORLD20:926-1102

# WORLD:87,147,...: obstacle_1..6; 7-14 are coincident at (0.22, 0.54) in the
# world file and pushed apart by ODE -- here they start on a ring around it.
_TRAIN_FIRST6 = [(-0.01, -1.0), (-1.15, -0.3), (-0.32, -0.12), (-0.85, 0.92), (0.94, 0.99), (0.65, 0.2)]
LAYOUT_TRAIN_14 = _TRAIN_FIRST6 + [
    (0.22 + 0.16 * math.cos(k * math.pi / 4), 0.54 + 0.16 * math.sin(k * math.pi / 4)) for k in range(8)
]
# WORLD20:86...1904 (ids 1-20)
LAYOUT_TEST_20 = [
    (-1.6, -1.3), (-1.0, -1.5), (-0.27, -1.47), (-1.28, -0.75), (-0.66, -0.86), (0.1, -0.81),
    (-1.63, 0.67), (-0.38, 0.45), (-1.46, 1.29), (-0.93, 0.76), (-0.48, 1.28), (0.056, 0.73),
    (0.310203, -1.50737), (0.422808, 0.415746), (0.676179, 1.21299), (-1.80625, -0.688364),
    (-2.00363, -1.5338), (-2.01729, 0.696956), (-2.05112, 1.57537), (0.537473, -0.824292),
]

# direction tables, unit = `speed` (simulate_{towards,crossing,ahead}_20.py:121-140)
_P, _M, _Z = 1.0, -1.0, 0.0
TABLE_TOWARDS_20 = [(_P, _P), (_P, _P), (_P, _P), (_P, _P), (_P, _Z), (_P, _P), (_P, _M), (_P, _M), (_P, _M),
                    (_P, _M), (_P, _M), (_P, _M), (_P, _P), (_P, _M), (_P, _M), (_P, _Z), (_P, _P), (_P, _Z),
                    (_P, _M), (_P, _P)]
TABLE_CROSSING_20 = [(_P, _P), (_Z, _P), (_Z, _P), (_P, _P), (_Z, _P), (_M, _P), (_Z, _M), (_M, _M), (_Z, _M),
                     (_Z, _M), (_Z, _M), (_M, _M), (_Z, _P), (_M, _M), (_M, _M), (_P, _P), (_P, _P), (_P, _M),
                     (_P, _M), (_Z, _P)]
TABLE_AHEAD_20 = [(_Z, _P), (_M, _P), (_M, _P), (_Z, _P), (_M, _Z), (_M, _Z), (_Z, _M), (_M, _Z), (_Z, _M),
                  (_M, _M), (_M, _M), (_M, _M), (_M, _P), (_M, _Z), (_M, _Z), (_P, _P), (_P, _P), (_P, _M),
                  (_P, _M), (_M, _P)]


def behavior_random(speed: float = 0.2, period_s: float = 1.5) -> Behavior:
    """simulate_crowd.py:101-126 (speed 0.2, one pass = 1.5 s) and the
    simulate_random_* family (other speeds / hold times)."""
    return Behavior(CN_BEHAVIOR_RANDOM, speed, period_s)


def behavior_table(table: Sequence[Sequence[float]], speed: float = 0.04, period_s: float = 0.5) -> Behavior:
    """simulate_{towards,crossing,ahead}_*.py: fixed velocities, re-sent continuously."""
    return Behavior(CN_BEHAVIOR_TABLE, speed, period_s, table=table)


def _hit_angle_inc_deg(n_samples: int) -> float:
    """UTL:113 `max_angle / (resolution - 1)` under Python 2 is INTEGER division
    (1 for 360 samples, 10 for 37); where that would be 0 (e.g. 721 samples) the
    true quotient is used."""
    q = 360 // (n_samples - 1)
    return float(q) if q >= 1 else 360.0 / (n_samples - 1)


def make_config(
    n_envs: int = 1,
    n_peds: int = 14,
    n_samples: int = 360,
    k_obstacles: int = 8,
    max_steps: int = 1000,
    room: Sequence[float] = ROOM_3M,
    start: Sequence[float] = (1.0, -1.0, 3.14),
    goal: Sequence[float] = (-1.0, 1.0),
    heading_offset: Sequence[float] = (0.75, -0.75),
    layout: Sequence[Sequence[float]] | None = None,
    behaviors: Sequence[Behavior] | None = None,
    layout_jitter: float = 0.0,
    collision_range: float = 0.12,
    seed: int = 1234,
    env_id_offset: int = 0,
    auto_reset: bool = False,
    topk_highest: bool = False,
    dt: float = 0.15,
    wheel_accel: float = 0.0,
    n_substeps: int = 1,
    env_original: bool = False,
    risk_faithful: bool = False,
) -> CnConfig:
    """Build a config; defaults are the reference's TRAINING world
    (CFG:1-18, WORLD, put_robot_in_world_training.launch:3-8)."""
    cfg = CnConfig()
    cfg.struct_size = C.sizeof(CnConfig)
    cfg.flags = (CN_FLAG_AUTO_RESET if auto_reset else 0) | (CN_FLAG_TOPK_HIGHEST if topk_highest else 0)
    if risk_faithful:
        if env_original:
            raise ValueError("risk_faithful has no meaning for the original environment (no K block)")
        cfg.flags |= CN_FLAG_RISK_FAITHFUL
    if env_original:
        # environment_stage_1_original.py: row = [ranges | heading, distance to the goal | x, y] (original:315-318), the
        # heading has no starting_pose term (original:246-260), an episode ends below 0.105 m (original:282)
        cfg.flags |= CN_FLAG_ENV_ORIGINAL
        k_obstacles, heading_offset = 0, (0.0, 0.0)
        if collision_range == 0.12:
            collision_range = 0.105
    cfg.n_envs, cfg.n_peds, cfg.n_samples, cfg.k_obstacles = n_envs, n_peds, n_samples, k_obstacles
    cfg.max_steps = max_steps
    cfg.env_id_offset = env_id_offset
    cfg.seed = seed
    cfg.dt = dt                                     # ENV:1201 (0.15 s sleep; ~0.19 s with the Python work around it)
    cfg.wheel_accel = wheel_accel                   # XACRO:70 (0 = instantaneous wheels, like FAKE:116-117)
    cfg.n_substeps = n_substeps                     # XACRO:65
    cfg.room_xmin, cfg.room_xmax, cfg.room_ymin, cfg.room_ymax = room
    cfg.start_x, cfg.start_y, cfg.start_yaw = start
    cfg.goal_x, cfg.goal_y = goal
    cfg.heading_off_x, cfg.heading_off_y = heading_offset   # CFG:15-18, ENV:223-224
    cfg.max_range = 0.6                             # CFG:7
    cfg.collision_range = collision_range           # CFG:8 (README test protocol: 0.0)
    cfg.sensor_min_range = 0.08                     # XACRO:164
    cfg.sensor_sweep = 6.28                         # XACRO:160 (not 2*pi)
    cfg.mount_x = -0.032                            # URDF:137
    cfg.hit_angle_inc_deg = _hit_angle_inc_deg(n_samples)
    cfg.ped_radius = 0.0505                         # WORLD:109
    cfg.robot_radius = 0.105                        # burger footprint, URDF:23-27,149-153
    cfg.cp_radius = 0.178                           # ENV:823
    cfg.waypoint_radius = 0.3                       # ENV:250
    cfg.goal_box = 0.20                             # ENV:1285,1303
    cfg.rep_strength = 0.5
    cfg.rep_range = 0.05
    cfg.rep_cutoff = 0.05
    cfg.layout_jitter = layout_jitter
    if layout is None:
        layout = LAYOUT_TRAIN_14
    if len(layout) < n_peds:
        raise ValueError("layout has %d poses, need %d" % (len(layout), n_peds))
    cfg.set_layout(list(layout)[:n_peds])
    if behaviors is None:
        behaviors = [behavior_random(0.2, 0.1 * n_peds + 0.1)]   # CROWD:48,144
    cfg.set_behaviors(behaviors)
    return cfg


def config_from_rosparams(params, **kw) -> CnConfig:
    """The reference's rosparam YAML (configs/turtlebot3_world.yaml, loaded under /turtlebot3 by
    launch/start_td3_training.launch:6-9 and read at ENV:71-90) -> a CnConfig.

    `params` is a path to such a YAML file or the already parsed mapping (with or without the `turtlebot3:` namespace
    level).  Keys used: scan_ranges -> n_samples, max_scan_range, min_scan_range -> collision_range (README.md:60-62
    sets it to 0.0 for evaluation), desired_pose -> goal, starting_pose -> the offset ENV:223-224 adds in the heading.
    The three discrete-action speeds (CFG:2-4) stay attributes of `Env`.  Everything else -- room, robot spawn pose,
    pedestrians, K -- comes from `kw` / make_config's defaults, as in the reference it comes from the world and launch files."""
    if isinstance(params, (str, bytes, os.PathLike)):
        import yaml
        with open(params) as fp:
            params = yaml.safe_load(fp)
    params = dict(params.get("turtlebot3", params))
    if "desired_pose" in params:
        kw.setdefault("goal", (float(params["desired_pose"]["x"]), float(params["desired_pose"]["y"])))
    if "starting_pose" in params:
        kw.setdefault("heading_offset", (float(params["starting_pose"]["x"]), float(params["starting_pose"]["y"])))
    if "scan_ranges" in params:
        kw.setdefault("n_samples", int(params["scan_ranges"]))
    if "min_scan_range" in params:
        kw.setdefault("collision_range", float(params["min_scan_range"]))
    cfg = make_config(**kw)
    if "max_scan_range" in params:
        cfg.max_range = float(params["max_scan_range"])
    return cfg


def realworld_layout_config(**kw) -> CnConfig:
    """The 370-wide row of the reference's physical-robot environment (environment_stage_1_nobonus_realworld.py:
    731-744): 359 ranges + 7 + ONE obstacle slot holding the object with the highest collision probability
    (realworld:674-678) -- i.e. K = 1 with the `highest` selection.  Values are rounded like the simulation
    environment's row (np.around, ENV:1042); the real-world script leaves the slot unrounded."""
    kw.setdefault("k_obstacles", 1)
    kw.setdefault("topk_highest", True)
    return make_config(**kw)


def test_world_20(n_envs: int = 1, behaviors: Sequence[Behavior] | None = None, **kw) -> CnConfig:
    """README test protocol (README.md:60-83): 5 m room, start (1, 0), goal (-2, 2)."""
    kw.setdefault("n_peds", 20)
    kw.setdefault("room", ROOM_5M)
    kw.setdefault("start", (1.0, 0.0, 3.14))
    kw.setdefault("goal", (-2.0, 2.0))
    kw.setdefault("heading_offset", (1.0, 0.0))
    kw.setdefault("layout", LAYOUT_TEST_20)
    if behaviors is None:
        behaviors = [behavior_random(0.04, 11.25)]               # simulate_random_20.py:111-119
    return make_config(n_envs=n_envs, behaviors=behaviors, **kw)


def shipped_actor_world(n_envs: int = 1, **kw) -> CnConfig:
    """The training world as the shipped `turtlebot3_top_8_obstacle` TD3 actor appears to have seen it.

    Driven in this simulator the checkpoint steers from the spawn pose to (-0.70 +- 0.05, 0.77) and idles there, i.e.
    its goal box was centred near (-0.7, 0.7), not the (-1, 1) of the committed YAML (CFG:10-13); and it was trained on
    Gazebo's acceleration-limited wheels (XACRO:65,70) at an effective control period of ~0.19 s (0.15 s sleep + the
    Python work around it: 5.5 steps/s in the logs).  With those three settings it reaches the goal in
    ~54 % of episodes here, against 58 % logged in Gazebo; see tests/test_actor_dropin.py."""
    kw.setdefault("goal", (-0.7, 0.7))
    kw.setdefault("dt", 0.19)
    kw.setdefault("wheel_accel", 1.0)
    kw.setdefault("n_substeps", 15)
    kw.setdefault("layout_jitter", 0.05)
    return make_config(n_envs=n_envs, **kw)


def _scatter_layout(n: int, room: Sequence[float], min_sep: float, seed: int,
                    keep_out: Iterable[Sequence[float]] = ()) -> list:
    """Seeded rejection sampling of n non-overlapping poses (BASELINE config 5)."""
    rng = np.random.default_rng(seed)
    pts: list = []
    lo_x, hi_x, lo_y, hi_y = room[0] + 0.1, room[1] - 0.1, room[2] + 0.1, room[3] - 0.1
    keep_out = [tuple(k) for k in keep_out]
    while len(pts) < n:
        p = (float(rng.uniform(lo_x, hi_x)), float(rng.uniform(lo_y, hi_y)))
        if all(math.hypot(p[0] - q[0], p[1] - q[1]) >= min_sep for q in pts) and \
                all(math.hypot(p[0] - k[0], p[1] - k[1]) >= k[2] for k in keep_out):
            pts.append(p)
    return pts


def baseline_config(index: int, n_envs: int | None = None, env_id_offset: int = 0,
                    auto_reset: bool = True, seed: int = 1234) -> CnConfig:
    """The five workloads of BASELINE.json `configs`."""
    if index == 0:      # c1: 1 env, 5 peds, 37 samples (36 rays), K=3, 3 m room
        ids = [0, 1, 2, 3, 5]
        return make_config(n_envs=n_envs or 1, n_peds=5, n_samples=37, k_obstacles=3,
                           layout=[LAYOUT_TRAIN_14[i] for i in ids], behaviors=[behavior_random(0.2, 0.6)],
                           auto_reset=auto_reset, seed=seed, env_id_offset=env_id_offset)
    if index in (1, 2):  # c2 / c3: 4096 / 16384 envs, 20 peds, 360 samples, K=8, 5 m room
        return test_world_20(n_envs=n_envs or (4096 if index == 1 else 16384), n_samples=360, k_obstacles=8,
                             behaviors=[behavior_random(0.2, 1.5)], layout_jitter=0.05,
                             auto_reset=auto_reset, seed=seed, env_id_offset=env_id_offset)
    if index == 3:      # c4: 65536 envs, mixed random / towards / crossing
        return test_world_20(n_envs=n_envs or 65536, n_samples=360, k_obstacles=8,
                             behaviors=[behavior_random(0.2, 1.5), behavior_table(TABLE_TOWARDS_20, 0.2),
                                        behavior_table(TABLE_CROSSING_20, 0.2)],
                             layout_jitter=0.05, auto_reset=auto_reset, seed=seed, env_id_offset=env_id_offset)
    if index == 4:      # c5: 16384 envs, 50 peds, 721 samples, K=16
        layout = _scatter_layout(50, ROOM_5M, 0.15, 50, keep_out=[(1.0, 0.0, 0.5)])
        return test_world_20(n_envs=n_envs or 16384, n_peds=50, n_samples=721, k_obstacles=16, layout=layout,
                             behaviors=[behavior_random(0.2, 1.5)], layout_jitter=0.0,
                             auto_reset=auto_reset, seed=seed, env_id_offset=env_id_offset)
    raise ValueError("baseline config index 0..4")


test_world_20.__test__ = False  # not a pytest test
