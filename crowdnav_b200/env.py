"""The reference's single-environment duck type on top of the batched CUDA env.

Mirrors ``Env`` of turtlebot3_rl_sim/src/environment_stage_1_nobonus.py so the
reference's drivers (start_td3_training.py:106-148, start_sac_training.py,
start_ddpg_training.py) and agents run unchanged:

    env = Env(action_dim=2, max_step=1000)
    state = env.reset()                       # np.ndarray [(R-1) + 7 + 4K]
    env.done = False                          # TD3DRV:116
    state, reward, done = env.step([v, w], step + 1, mode="continuous")
    success, failure = env.get_episode_status()
    env.get_social_safety_violation_status(step + 1)

There is no ROS, no Gazebo and no sleep: one call = one kernel launch on the
GPU plus the host copies of a 2-float action and one observation row.
"""
from __future__ import annotations

import numpy as np

from .config import CnConfig, make_config
from .vec_env import CrowdNavVecEnv


class Env:
    # CFG:2-4 discrete action set (ENV:1165-1177)
    linear_forward_speed = 0.5
    linear_turn_speed = 0.05
    angular_speed = 0.3

    def __init__(self, action_dim: int = 2, max_step: int = 200, config: CnConfig | None = None,
                 device: int | None = None):
        cfg = (config.copy() if config is not None else make_config())
        cfg.n_envs = 1
        cfg.max_steps = max_step
        cfg.flags &= ~1          # the driver resets explicitly (TD3DRV:113)
        self.action_dim = action_dim
        self.max_steps = max_step
        self.k_obstacle_count = cfg.k_obstacles            # ENV:55
        self.scan_ranges = cfg.n_samples
        self.max_scan_range = cfg.max_range
        self.min_scan_range = cfg.collision_range
        self._venv = CrowdNavVecEnv(cfg, device=device)
        self._done = False
        self.episode_success = False
        self.episode_failure = False

    # `done` is written by the drivers after reset (TD3DRV:116)
    @property
    def done(self) -> bool:
        return self._done

    @done.setter
    def done(self, value: bool) -> None:
        self._done = bool(value)
        if not self._done:
            self._venv.clear_done()

    def reset(self) -> np.ndarray:
        """ENV:1227-1263."""
        obs = self._venv.reset()
        return obs[0].detach().cpu().numpy().astype(np.float64)

    def step(self, action, step_counter, mode: str = "discrete"):
        """ENV:1164-1225.  `step_counter` is the driver's 1-based step index; the
        kernel keeps the same count per world."""
        if mode == "discrete":
            if action == 0:
                v, w = self.linear_forward_speed, 0.0
            elif action == 1:
                v, w = self.linear_turn_speed, self.angular_speed
            elif action == 2:
                v, w = self.linear_turn_speed, -1.0 * self.angular_speed
            else:
                raise ValueError("discrete action must be 0, 1 or 2")
        else:
            v, w = float(action[0]), float(action[1])
        obs, reward, done = self._venv.step_host(np.array([[v, w]], dtype=np.float32))
        self._done = bool(done[0])
        if self._done:
            success = int(self._venv.counters()[0, 0].item()) == 1
            self.episode_success, self.episode_failure = success, not success     # ENV:1148-1159
        r = float(reward[0])
        return obs[0].astype(np.float64), (int(r) if r == int(r) else r), self._done

    def get_episode_status(self):
        """ENV:1265-1267."""
        return self.episode_success, self.episode_failure

    def _counts(self):
        c = self._venv.counters()[0].tolist()
        return c[1], c[2], c[3]

    def get_social_safety_violation_status(self, step):
        """ENV:1269-1275 (divides by zero, like the reference, if no obstacle was ever seen)."""
        _, social, present = self._counts()
        return 1.0 - ((social * 1.0) / present)

    def get_ego_safety_violation_status(self, step):
        """ENV:1277-1283."""
        ego, _, present = self._counts()
        return 1.0 - ((ego * 1.0) / present)

    def shutdown(self) -> None:
        """ENV:170-173: publish a zero twist.  Nothing to stop here."""
        return None


class EnvOriginal(Env):
    """``Env`` of turtlebot3_rl_sim/src/environment_stage_1_original.py (the reference's first environment: DQN /
    tabular drivers, the 363-wide TD3 / DDPG ``trajectory_test`` checkpoints): row = [359 ranges | heading, distance
    to the goal | x, y], reward = progress terms + terminal (original:324-410), discrete actions 0/1/2 =
    (0.22, 0) / (0.22, +2) / (0.22, -2) (original:412-425)."""

    def __init__(self, action_dim: int = 2, max_step: int = 200, config: CnConfig | None = None,
                 device: int | None = None):
        cfg = config.copy() if config is not None else make_config(env_original=True)
        if not cfg.flags & 4:
            raise ValueError("EnvOriginal needs a config made with env_original=True")
        super().__init__(action_dim, max_step, cfg, device)

    def step(self, action, step_counter, mode: str = "discrete"):
        if mode == "discrete":
            if action not in (0, 1, 2):
                raise ValueError("discrete action must be 0, 1 or 2")
            action = [0.22, (0.0, 2.0, -2.0)[action]]
        return super().step(action, step_counter, mode="continuous")

    def get_odometry_data(self):
        """original:494-496: [round(x, 3), round(y, 3), yaw] of the last get_state."""
        row = self._venv.obs[0].detach().cpu().numpy()
        blob = self._venv.get_state_blob()
        yaw = float(np.float32(np.int32(blob[16 + 2])) * np.float32(1.4629180792671596e-09))
        return [float(row[-2]), float(row[-1]), yaw]
