"""The reference's single-env duck type (crowdnav_b200.env.Env) driven the way the reference's
training driver drives it (start_td3_training.py:106-148), checked against the oracle."""
import numpy as np
import pytest

from crowdnav_b200.config import make_config
from parity_util import random_actions

pytestmark = pytest.mark.gpu


def test_env_duck_type_follows_the_reference_driver_protocol():
    from crowdnav_b200.env import Env
    from oracle.oracle import OracleEnv
    cfg = make_config(layout_jitter=0.05)
    env = Env(action_dim=2, max_step=60, config=cfg)
    ocfg = cfg.copy()
    ocfg.max_steps = 60
    o = OracleEnv(ocfg)
    rng = np.random.default_rng(0)
    for ep in range(4):
        observation = env.reset()                      # TD3DRV:113
        o.reset()
        env.done = False                               # TD3DRV:116
        o.clear_done()
        assert observation.shape == (398,) and observation.dtype == np.float64
        assert np.array_equal(np.float32(observation), o.obs[0])
        for step in range(60):
            a = random_actions(rng, 1)
            state, reward, done = env.step(a[0].tolist(), step + 1, mode="continuous")   # TD3DRV:125
            oo, orr, od = o.step(a)
            assert np.array_equal(np.float32(state), oo[0]) and reward == orr[0] and done == bool(od[0])
            assert isinstance(done, bool)
            success, failure = env.get_episode_status()                                    # TD3DRV:126
            if done:
                assert success != failure
                c = o.counters()[0]
                if c[3] > 0:
                    assert env.get_social_safety_violation_status(step + 1) == 1.0 - c[2] / c[3]   # TD3DRV:144
                    assert env.get_ego_safety_violation_status(step + 1) == 1.0 - c[1] / c[3]
                else:
                    with pytest.raises(ZeroDivisionError):      # ENV:1272 divides by zero too
                        env.get_social_safety_violation_status(step + 1)
                break
        assert done, "max_step=60 must end the episode (ENV:1021)"
    env.shutdown()


def test_env_discrete_mode_actions():
    """ENV:1165-1177 / CFG:2-4: 0 = forward 0.5, 1 / 2 = turn 0.05 with +-0.3 rad/s."""
    from crowdnav_b200.env import Env
    from oracle.oracle import OracleEnv
    cfg = make_config()
    env = Env(action_dim=3, max_step=50, config=cfg)
    ocfg = cfg.copy()
    ocfg.max_steps = 50
    o = OracleEnv(ocfg)
    env.reset()
    o.reset()
    env.done = False
    table = {0: (0.5, 0.0), 1: (0.05, 0.3), 2: (0.05, -0.3)}
    for step, act in enumerate([1, 1, 2, 0, 2, 1]):
        s, r, d = env.step(act, step + 1)               # mode defaults to "discrete" like the reference
        oo, orr, od = o.step(np.array([table[act]], dtype=np.float32))
        assert np.array_equal(np.float32(s), oo[0]) and r == orr[0]
        if d:
            break
