"""Checkpoint drop-in (fixture row 1): the reference's shipped K=8 TD3 actor
(models/td3/turtlebot3_top_8_obstacle/td3_actor_model_ep2500.pt, committed as an npz of its tensors) consumes this
environment's observations unchanged and reaches the goal at a rate in the ballpark of its Gazebo logs (0.58)."""
import os

import numpy as np
import pytest
import torch

from crowdnav_b200.config import make_config, shipped_actor_world
from crowdnav_b200.rollout import ReplayRing, TD3Learner, collect, load_reference_actor

ACTOR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "td3_actor_k8_ep2500.npz")


def test_actor_input_width_is_the_observation_width():
    actor = load_reference_actor(ACTOR)
    assert actor.linear1.in_features == make_config(k_obstacles=8).obs_dim == 398
    a = actor(torch.zeros(3, 398))
    assert a.shape == (3, 2) and (a[:, 0] >= 0).all() and (a[:, 0] <= 0.22).all() and (a[:, 1].abs() <= 2.0).all()


def test_shipped_actor_reaches_goal_on_the_oracle():
    """CPU: 192 worlds on the oracle.  Success in the ballpark of the logged 0.58; failures are collisions."""
    from oracle.oracle import OracleEnv
    actor = load_reference_actor(ACTOR)
    E = 192
    env = OracleEnv(shipped_actor_world(n_envs=E, max_steps=600))
    obs = env.reset().copy()
    alive = np.ones(E, bool)
    succ = 0
    for t in range(600):
        with torch.no_grad():
            a = actor(torch.from_numpy(obs)).numpy().astype(np.float32)
        o, r, d = env.step(a)
        ended = alive & (d > 0)
        succ += int(env.counters()[ended, 0].sum())
        alive &= ~ended
        obs = o.copy()
        if not alive.any():
            break
    rate = succ / E
    assert 0.35 <= rate <= 0.85, "success rate %.2f is not in the ballpark of the logged 0.58" % rate


@pytest.mark.gpu
def test_shipped_actor_rollout_on_gpu_and_td3_update():
    """GPU: batched rollout of the shipped actor (4096 worlds), device replay ring, a few TD3 updates."""
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    dev = torch.device("cuda", 0)
    actor = load_reference_actor(ACTOR, dev)
    env = CrowdNavVecEnv(shipped_actor_world(n_envs=4096, max_steps=600, auto_reset=True), device=0)
    env.reset()
    replay = ReplayRing(200_000, env.D, dev)
    stats = collect(env, actor, 400, sigma=0.0, replay=replay)
    assert stats["episodes"] > 3000 and stats["censored"] <= 4096
    assert 0.35 <= stats["success_rate"] <= 0.85, stats
    assert len(replay) == 200_000
    s, a, r, s2, d = replay.sample(128)
    assert s.shape == (128, 398) and set(torch.unique(d).tolist()) <= {0.0, 1.0}
    learner = TD3Learner(env.D, dev)
    losses = [learner.learn(replay.sample(128)) for _ in range(6)]
    assert all(np.isfinite(float(v)) for l in losses for v in l.values()) and any("actor" in l for l in losses)
    # exploration (TD3:67-78) keeps actions inside the box
    noisy = collect(env, actor, 20, sigma=1.0)
    assert noisy["episodes"] >= 0


@pytest.mark.gpu
def test_graphed_collector_equals_the_eager_loop():
    """The rollout loop captured as a CUDA graph (GraphedCollector) does what the eager loop does: same worlds, greedy
    actor, 3 warm-up iterations + 40 replays vs 43 eager steps -> identical state, replay rows and episode statistics;
    and a graph with the TD3 update inside runs."""
    from crowdnav_b200.rollout import GraphedCollector
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    dev = torch.device("cuda", 0)
    actor = load_reference_actor(ACTOR, dev)
    cfg = shipped_actor_world(n_envs=512, max_steps=60, auto_reset=True)
    e1, e2 = CrowdNavVecEnv(cfg, device=0), CrowdNavVecEnv(cfg, device=0)
    e1.reset(); e2.reset()
    r1, r2 = ReplayRing(50_000, e1.D, dev), ReplayRing(50_000, e2.D, dev)
    eager = collect(e1, actor, 43, sigma=0.0, replay=r1)
    gc = GraphedCollector(e2, actor, sigma=0.0, replay=r2)
    assert gc.run(40) == 40
    torch.cuda.synchronize()
    got = gc.stats()
    assert np.array_equal(e1.get_state_blob(), e2.get_state_blob())
    assert len(r1) == len(r2) and torch.equal(r1.state[:len(r1)], r2.state[:len(r2)]) and torch.equal(r1.reward, r2.reward)
    for k in ("episodes", "successes", "mean_return", "mean_length"):
        assert eager[k] == got[k], (k, eager[k], got[k])
    learner = TD3Learner(e2.D, dev, capturable=True)
    gl = GraphedCollector(e2, learner.actor, sigma=1.0, replay=r2, learner=learner, batch_size=128)
    before = [p.detach().clone() for p in learner.critic1.parameters()]
    assert gl.run(10) == 10
    torch.cuda.synchronize()
    assert any(not torch.equal(a, b) for a, b in zip(before, learner.critic1.parameters()))
    assert all(torch.isfinite(p).all() for p in learner.actor.parameters())
