"""GPU-vs-oracle parity of the `risk_faithful` block (CN_FLAG_RISK_FAITHFUL), through the C ABI.

The oracle (oracle/cn_oracle_faithful.c) is pinned to the reference itself by tests/test_faithful.py; here
cn_faithful_kernel (one warp per world, float64) must reproduce it BIT FOR BIT: observation rows (K block
included), the whole state blob with the tracker plane, rewards, done flags and the safety counters.
"""
import numpy as np
import pytest

from crowdnav_b200.config import CN_FLAG_RISK_FAITHFUL, baseline_config, make_config
from parity_util import bits_equal, random_actions
from test_gpu_parity import _rollout

pytestmark = pytest.mark.gpu


def _faithful(cfg):
    cfg = cfg.copy()
    cfg.flags |= CN_FLAG_RISK_FAITHFUL
    return cfg


def _forward(rng, E):
    a = np.empty((E, 2), dtype=np.float32)
    a[:, 0] = 0.22
    a[:, 1] = rng.uniform(-0.5, 0.5, E)
    return a


def test_c1_200_steps():
    _rollout(_faithful(baseline_config(0, n_envs=1, auto_reset=True)), 200)


def test_training_world():
    cfg = _faithful(make_config(n_envs=64, auto_reset=True, layout_jitter=0.05, max_steps=120))
    assert _rollout(cfg, 150, seed=2, action_fn=lambda rng, t: _forward(rng, 64)) > 0


def test_c2_shape_ragged_batch():
    # 203 worlds: the last CTA of cn_faithful_kernel (4 worlds each) is partly empty
    cfg = _faithful(baseline_config(1, n_envs=203, auto_reset=True))
    assert _rollout(cfg, 120, seed=3, action_fn=lambda rng, t: _forward(rng, 203)) > 0


def test_c2_random_policy():
    _rollout(_faithful(baseline_config(1, n_envs=512, auto_reset=True)), 60, seed=4, check_every=5)


def test_c5_shape():
    cfg = _faithful(baseline_config(4, n_envs=48, auto_reset=True))
    _rollout(cfg, 40, seed=5, action_fn=lambda rng, t: _forward(rng, 48))


def test_k1_highest():
    cfg = _faithful(baseline_config(1, n_envs=96, auto_reset=True))
    cfg.flags |= 2
    cfg.k_obstacles = 1
    _rollout(cfg, 80, seed=6, action_fn=lambda rng, t: _forward(rng, 96))


def test_full_c2_without_debug_taps_masked_reset_and_blob_round_trip():
    """configs[1] at full size, ranges handed to the block through the handle's own buffer (no debug taps);
    masked reset keeps the other worlds' trackers; a blob written back resumes exactly."""
    import torch
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    from oracle.oracle import OracleEnv
    cfg = _faithful(baseline_config(1, auto_reset=False))
    E = cfg.n_envs
    g, o = CrowdNavVecEnv(cfg, device=0), OracleEnv(cfg)
    rng = np.random.default_rng(7)
    g.reset()
    o.reset()
    for t in range(30):
        a = _forward(rng, E)
        g.step(torch.from_numpy(a).cuda())
        o.step(a)
    torch.cuda.synchronize()
    assert bits_equal(g.obs.cpu().numpy(), o.obs) and bits_equal(g.get_state_blob(), o.blob)
    assert o.trk()[:, 0].max() >= 2                      # several objects tracked somewhere
    mask = (rng.uniform(size=E) < 0.3).astype(np.uint8)
    g.reset(torch.from_numpy(mask).cuda())
    o.reset(mask)
    g.clear_done()
    o.clear_done()
    blob = g.get_state_blob()
    assert bits_equal(blob, o.blob)
    g2 = CrowdNavVecEnv(cfg, device=0)
    g2.reset()
    g2.set_state_blob(blob)
    for t in range(10):
        a = random_actions(rng, E)
        g.step(torch.from_numpy(a).cuda())
        g2.step(torch.from_numpy(a).cuda())
        o.step(a)
    torch.cuda.synchronize()
    assert bits_equal(g.obs.cpu().numpy(), o.obs) and bits_equal(g2.obs.cpu().numpy(), o.obs)
    assert bits_equal(g.get_state_blob(), o.blob) and bits_equal(g2.get_state_blob(), o.blob)
    assert bits_equal(g.counters().cpu().numpy(), o.counters())
    assert g.launch_count == 2 * (1 + 30 + 1 + 10) + 1 + 1   # step kernel + cn_faithful_kernel per call; clear_done; counters
    g.close()
    g2.close()


def test_fused_gather_is_refused():
    import ctypes as C
    import torch
    from crowdnav_b200 import _lib
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    cfg = _faithful(baseline_config(1, n_envs=16, auto_reset=True))
    g = CrowdNavVecEnv(cfg, device=0)
    g.reset()
    peer = torch.zeros((16, cfg.obs_dim), device="cuda")
    ptrs = (C.c_void_p * 1)(peer.data_ptr())
    a = torch.zeros((16, 2), device="cuda")
    rc = g._L.cn_step_gather(g._h, C.c_void_p(a.data_ptr()), C.c_void_p(g.obs.data_ptr()), ptrs, 1,
                             C.c_void_p(g.reward.data_ptr()), C.c_void_p(g.done.data_ptr()), g._stream())
    assert rc == -4 and b"RISK_FAITHFUL" in _lib.load().cn_last_error()
    g.close()


def test_duck_type_and_evaluation_harness_with_the_faithful_block():
    """The single-env `Env` duck type and the README evaluation protocol run on the reference's own perception
    block: same row width, counters read from the tracker record, scores like ENV:1269-1283."""
    import os
    from crowdnav_b200.env import Env
    from crowdnav_b200.evaluate import evaluate, scenario_config, summarize
    from crowdnav_b200.rollout import load_reference_actor
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    from oracle.oracle import OracleEnv
    cfg = make_config(risk_faithful=True)
    env = Env(action_dim=2, max_step=60, config=cfg)
    ocfg = cfg.copy()
    ocfg.n_envs, ocfg.max_steps = 1, 60
    ocfg.flags &= ~1
    o = OracleEnv(ocfg)
    s = env.reset()
    o.reset()
    env.done = False
    assert s.shape == (398,) and np.array_equal(s.astype(np.float32), o.obs[0])
    for t in range(40):
        a = [0.2, 0.3 if t % 7 else -0.4]
        s, r, d = env.step(a, t + 1, mode="continuous")
        oo, orr, od = o.step(np.array([a], dtype=np.float32))
        assert np.array_equal(s.astype(np.float32), oo[0]) and r == orr[0] and d == bool(od[0])
        if d:
            break
    assert list(env._counts()) == [int(v) for v in o.counters()[0, 1:]]
    actor = load_reference_actor(os.path.join(os.path.dirname(__file__), "golden", "td3_actor_k8_ep2500.npz"), "cuda")
    venv = CrowdNavVecEnv(scenario_config("crossing", 8, n_envs=256, max_steps=300, risk_faithful=True), device=0)
    rows = evaluate(venv, actor, 300)
    sm = summarize(rows)
    assert len(rows) == 512 and sm["mean_steps"] <= 300 and sm["ego_safety"] <= 1.0 and sm["social_safety"] <= 1.0
