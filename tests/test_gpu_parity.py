"""GPU-vs-oracle parity, through the C ABI (CrowdNavVecEnv -> libcrowdnav.so).

Bar: BIT-EXACT.  State blobs (integer-grid poses, fp32 fields), observation
rows, raw LiDAR ranges, hit ids, rewards and done flags must equal the CPU
oracle's on the same seeded inputs -- stronger than the 1e-4 fp32 tolerance
BASELINE.json allows for ranges / reward (tolerance used here: 0).
"""
import numpy as np
import pytest

from crowdnav_b200.config import (TABLE_TOWARDS_20, baseline_config, behavior_random, behavior_table, make_config,
                                  test_world_20)
from parity_util import bits_equal, describe_blob_diff, describe_obs_diff, random_actions

pytestmark = pytest.mark.gpu


def _mk(cfg):
    import torch
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    from oracle.oracle import OracleEnv
    g = CrowdNavVecEnv(cfg, device=0)
    g.enable_debug_taps()
    o = OracleEnv(cfg, debug=True)
    return torch, g, o


def _check(cfg, g, o, tag):
    import torch
    torch.cuda.synchronize()
    gb, ob = g.get_state_blob(), o.blob
    assert bits_equal(gb, ob), "%s: state blob differs\n%s" % (tag, describe_blob_diff(cfg, gb, ob))
    go = g.obs.cpu().numpy()
    assert bits_equal(go, o.obs), "%s: observation differs\n%s" % (tag, describe_obs_diff(cfg, go, o.obs))


def _rollout(cfg, steps, seed=0, check_every=1, action_fn=None):
    torch, g, o = _mk(cfg)
    rng = np.random.default_rng(seed)
    g.reset()
    o.reset()
    assert bits_equal(g.debug_hit_ids.cpu().numpy(), o.hit_ids), "reset: hit ids differ"
    assert bits_equal(g.debug_ranges.cpu().numpy(), o.ranges), "reset: raw ranges differ"
    _check(cfg, g, o, "reset")
    n_done = 0
    for t in range(steps):
        a = action_fn(rng, t) if action_fn else random_actions(rng, cfg.n_envs)
        _, gr, gd = g.step(torch.from_numpy(a).cuda())
        _, orr, od = o.step(a)
        if (t + 1) % check_every == 0 or t == steps - 1:
            tag = "step %d" % (t + 1)
            assert bits_equal(g.debug_hit_ids.cpu().numpy(), o.hit_ids), tag + ": hit ids differ"
            assert bits_equal(g.debug_ranges.cpu().numpy(), o.ranges), tag + ": raw ranges differ"
            _check(cfg, g, o, tag)
            assert bits_equal(gr.cpu().numpy(), orr), tag + ": reward differs"
            assert bits_equal(gd.cpu().numpy(), od), tag + ": done differs"
        n_done += int(od.sum())
    assert bits_equal(g.counters().cpu().numpy(), o.counters()), "counters differ"
    g.close()
    return n_done


def test_c1_single_env_200_steps():
    """BASELINE config 1: 1 env, 5 pedestrians, 36 rays, K=3, random policy, 200 steps."""
    cfg = baseline_config(0, auto_reset=False)
    _rollout(cfg, 200)


def test_c1_auto_reset_many_envs():
    cfg = baseline_config(0, n_envs=333, auto_reset=True)
    n_done = _rollout(cfg, 300, seed=1)
    assert n_done > 0, "rollout never finished an episode: auto-reset path not exercised"


def test_training_world_14_peds():
    """The reference's training world (3 m room, 14 pedestrians, 360 samples, K=8)."""
    cfg = make_config(n_envs=200, auto_reset=True, layout_jitter=0.05)
    n_done = _rollout(cfg, 250, seed=2)
    assert n_done > 0


def test_c2_shape_20_peds_360_rays():
    """BASELINE config 2 shape at a size the oracle steps in seconds."""
    cfg = baseline_config(1, n_envs=1024)
    n_done = _rollout(cfg, 200, seed=3, check_every=10)
    assert n_done > 0


def test_c4_mixed_behaviours_and_sharding_offset():
    """Config 4 shape: behaviours by env_id % 3, and a non-zero global env id offset."""
    cfg = baseline_config(3, n_envs=600, env_id_offset=8192)
    _rollout(cfg, 120, seed=4, check_every=10)


def test_c5_50_peds_720_rays_k16():
    """BASELINE config 5 shape: two pedestrians per lane, 720 rays, K=16."""
    cfg = baseline_config(4, n_envs=160)
    _rollout(cfg, 120, seed=5, check_every=10)


def test_partial_tile_and_odd_obs_dim():
    """E not a multiple of the CTA tile, odd observation width (plain-store path)."""
    cfg = make_config(n_envs=37, n_peds=7, n_samples=38, k_obstacles=2, auto_reset=True, layout_jitter=0.1)
    _rollout(cfg, 150, seed=6)


def test_crowded_contacts():
    """Dense crowd in a small area: the contact stand-in (repulsion) path is active."""
    layout = [(-0.3 + 0.11 * (i % 6), -0.3 + 0.11 * (i // 6)) for i in range(30)]
    cfg = make_config(n_envs=64, n_peds=30, layout=layout, behaviors=[behavior_random(0.2, 0.6)], auto_reset=True,
                      layout_jitter=0.02, start=(1.0, -1.0, 3.14))
    _rollout(cfg, 100, seed=7)


def test_drive_at_pedestrians_and_walls():
    """Straight driving so collisions (range < 0.12) and the sensor-minimum clamp occur."""
    cfg = test_world_20(n_envs=256, behaviors=[behavior_table(TABLE_TOWARDS_20, 0.2)], auto_reset=True,
                        layout_jitter=0.3, start=(0.9, 0.0, 3.14))

    def act(rng, t):
        a = np.zeros((256, 2), dtype=np.float32)
        a[:, 0] = 0.22
        a[:, 1] = rng.uniform(-0.3, 0.3, 256)
        return a
    n_done = _rollout(cfg, 200, seed=8, action_fn=act)
    assert n_done > 20


@pytest.mark.parametrize("kernel", ["flat", "warp"])
def test_eval_protocol_body_contact(kernel, monkeypatch):
    """README test protocol (`min_scan_range 0.0`): the robot's body is stopped by walls and pedestrians instead of
    the LiDAR threshold ending the episode (robot_contact_step); wheel ramp + 15 sub-steps so that a control period is
    partly blocked.  Both kernels, bit for bit against the oracle."""
    from crowdnav_b200.evaluate import scenario_config
    monkeypatch.setenv("CN_KERNEL", kernel)
    cfg = scenario_config("towards", 20, n_envs=200, max_steps=150, wheel_accel=1.0, n_substeps=15, layout_jitter=0.3,
                          start=(0.9, 0.0, 3.14))

    def act(rng, t):
        a = np.zeros((200, 2), dtype=np.float32)
        a[:, 0] = 0.22
        a[:, 1] = rng.uniform(-0.4, 0.4, 200)
        return a
    _rollout(cfg, 220, seed=31, action_fn=act)
    cfg = make_config(n_envs=64, collision_range=0.0, auto_reset=True, layout_jitter=0.05, max_steps=120)   # 3 m room: walls
    _rollout(cfg, 150, seed=32, action_fn=lambda rng, t: np.stack([np.full(64, 0.22), rng.uniform(-0.2, 0.2, 64)], 1).astype(np.float32))


def test_nonfinite_and_out_of_range_actions():
    cfg = baseline_config(0, n_envs=32, auto_reset=True)

    def act(rng, t):
        a = random_actions(rng, 32)
        a[0, 0] = np.nan
        a[1, 1] = np.inf
        a[2, 0] = 50.0
        a[3, 1] = -50.0
        return a
    _rollout(cfg, 40, seed=9, action_fn=act)


def test_masked_reset_and_blob_roundtrip():
    cfg = baseline_config(1, n_envs=100, auto_reset=False)
    torch, g, o = _mk(cfg)
    rng = np.random.default_rng(10)
    g.reset()
    o.reset()
    for t in range(30):
        a = random_actions(rng, 100)
        g.step(torch.from_numpy(a).cuda())
        o.step(a)
    mask = (rng.uniform(size=100) < 0.3).astype(np.uint8)
    before = g.obs.cpu().numpy().copy()
    g.reset(torch.from_numpy(mask))
    o.reset(mask)
    _check(cfg, g, o, "masked reset")
    after = g.obs.cpu().numpy()
    assert bits_equal(after[mask == 0], before[mask == 0]), "masked reset touched unmasked rows"
    # blob round trip: restore an earlier snapshot and replay
    snap = g.get_state_blob().copy()
    acts = [random_actions(rng, 100) for _ in range(10)]
    outs = []
    for a in acts:
        ob, r, d = g.step(torch.from_numpy(a).cuda())
        outs.append((ob.cpu().numpy().copy(), r.cpu().numpy().copy(), d.cpu().numpy().copy()))
    g.set_state_blob(snap)
    for a, (ob0, r0, d0) in zip(acts, outs):
        ob, r, d = g.step(torch.from_numpy(a).cuda())
        assert bits_equal(ob.cpu().numpy(), ob0) and bits_equal(r.cpu().numpy(), r0) and bits_equal(d.cpu().numpy(), d0)
    g.close()


def test_full_size_c2_4096_envs():
    """BASELINE configs[1] at its full size: 4096 envs x 60 steps, bit-exact at every 20th step."""
    _rollout(baseline_config(1), 60, seed=21, check_every=20)


def test_full_size_c3_16384_envs():
    """BASELINE configs[2] at its full size: 16384 envs x 20 steps."""
    _rollout(baseline_config(2), 20, seed=22, check_every=20)


def test_sharding_invariance_on_device():
    """Envs [lo, hi) of a 4096-env launch == a (hi-lo)-env launch with env_id_offset = lo (what each rank of the
    8-GPU config computes): concatenated results are independent of the number of GPUs."""
    import torch
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    E, lo, hi = 4096, 1024, 1536
    full = CrowdNavVecEnv(baseline_config(3, n_envs=E), device=0)
    part = CrowdNavVecEnv(baseline_config(3, n_envs=hi - lo, env_id_offset=lo), device=0)
    rng = np.random.default_rng(23)
    full.reset()
    part.reset()
    for t in range(50):
        a = torch.from_numpy(random_actions(rng, E)).cuda()
        fo, fr, fd = full.step(a)
        po, pr, pd = part.step(a[lo:hi].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(fo[lo:hi], po) and torch.equal(fr[lo:hi], pr) and torch.equal(fd[lo:hi], pd)
    assert bits_equal(full.get_state_blob()[16 + lo * 16:16 + hi * 16], part.get_state_blob()[16:16 + (hi - lo) * 16])
    full.close()
    part.close()


def test_wheel_ramp_and_substeps():
    """libgazebo_ros_diff_drive-like dynamics: 15 kinematic sub-steps, 1 m/s^2 wheel ramp, 0.19 s period."""
    from crowdnav_b200.config import shipped_actor_world
    cfg = shipped_actor_world(n_envs=300, auto_reset=True, max_steps=200)
    n_done = _rollout(cfg, 150, seed=31)
    assert n_done > 0


def test_cuda_graph_capture_of_cn_step():
    """cn_step enqueues on the caller's stream with no hidden synchronisation: it can be captured into a CUDA graph
    and replayed (what a launch-bound rollout loop does); results equal the oracle's."""
    import torch
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    from oracle.oracle import OracleEnv
    cfg = baseline_config(1, n_envs=512)
    g, o = CrowdNavVecEnv(cfg, device=0), OracleEnv(cfg)
    g.reset()
    o.reset()
    rng = np.random.default_rng(41)
    act = torch.zeros((512, 2), device="cuda")
    a0 = random_actions(rng, 512)
    act.copy_(torch.from_numpy(a0))
    g.step(act)                      # warm-up launch outside capture (sets the kernel attributes)
    o.step(a0)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g.step(act)
    for _ in range(25):
        a = random_actions(rng, 512)
        act.copy_(torch.from_numpy(a))
        graph.replay()
        o.step(a)
    torch.cuda.synchronize()
    assert bits_equal(g.obs.cpu().numpy(), o.obs) and bits_equal(g.get_state_blob(), o.blob)
    g.close()


@pytest.mark.parametrize("kernel", ["warp", "flat"])
def test_both_step_kernels_against_the_oracle(kernel, monkeypatch):
    """The library carries two independent step kernels (cn_flat.cu: compacted work lists, the default;
    cn_step.cu: one warp per world, CN_KERNEL=warp).  Each must match the oracle bit for bit on a rollout that
    exercises auto-reset, contacts, walls and the K block -- and therefore each other."""
    monkeypatch.setenv("CN_KERNEL", kernel)
    cfg = make_config(n_envs=150, auto_reset=True, layout_jitter=0.05)
    torch, g, o = _mk(cfg)
    assert g.kernel_name == {"warp": "cn_env_kernel", "flat": "cn_flat_kernel"}[kernel]
    g.close()
    n_done = _rollout(cfg, 120, seed=51)
    assert n_done > 0
    n_done = _rollout(baseline_config(4, n_envs=96, auto_reset=True), 40, seed=52)     # 50 pedestrians, 720 rays, K = 16


@pytest.mark.parametrize("direct", ["1", "0"])
@pytest.mark.parametrize("tile", ["2,256", "5,256", "16,256", "6,128", "20,512", "19,256", "28,384", "32,512"])
def test_flat_kernel_tile_shapes(tile, direct, monkeypatch):
    """Tile size / CTA size are launch parameters of the flat kernel, not part of the result: odd tiles (rows leave by
    plain stores), ragged last tiles, several pedestrian passes per thread, group-list overflow (all worlds start next
    to two walls in the 3 m room) -- all bit-exact.  direct = 1: plain steps write their rows straight to global memory
    (the default, DIRECT instance of cn_flat_kernel); 0: the staged instance every fused-gather entry point uses."""
    if direct == "0" and int(tile.split(",")[0]) > 20:
        pytest.skip("the staged layout of this tile does not fit shared memory")
    monkeypatch.setenv("CN_KERNEL", "flat")
    monkeypatch.setenv("CN_FLAT_TILE", tile)
    monkeypatch.setenv("CN_FLAT_DIRECT", direct)
    cfg = make_config(n_envs=77, auto_reset=True, layout_jitter=0.05)
    torch, g, o = _mk(cfg)
    assert g.kernel_tile == int(tile.split(",")[0])
    g.close()
    _rollout(cfg, 60, seed=53)
    monkeypatch.setenv("CN_FLAT_STORE", "plain")
    _rollout(baseline_config(1, n_envs=130, auto_reset=True), 30, seed=54)


@pytest.mark.parametrize("off", [1, 2, 3])
def test_direct_rows_into_an_unaligned_buffer(off):
    """The direct-rows step kernel writes the no-return fill and the owned rays straight into the caller's buffer:
    a row block that starts 4 / 8 / 12 bytes off a 16-byte boundary (a slice of a larger tensor), with an odd number
    of worlds, must come out the same as the oracle's, and the bytes around the block must stay untouched."""
    import torch
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    from oracle.oracle import OracleEnv
    cfg = make_config(n_envs=53, auto_reset=True, layout_jitter=0.05)
    E, D = cfg.n_envs, cfg.obs_dim
    big = torch.full((E * D + 64,), -7.0, dtype=torch.float32, device="cuda")
    view = big[off:off + E * D].view(E, D)
    g, o = CrowdNavVecEnv(cfg, device=0, obs_out=view), OracleEnv(cfg)
    g.reset(); o.reset()
    rng = np.random.default_rng(81 + off)
    for t in range(25):
        a = random_actions(rng, E)
        obs, rew, done = g.step(torch.from_numpy(a).cuda())
        o_obs, o_rew, o_done = o.step(a)
        assert obs.data_ptr() == view.data_ptr()
        got = view.cpu().numpy()
        assert bits_equal(got, o_obs), describe_obs_diff(cfg, got, o_obs)
        assert np.array_equal(rew.cpu().numpy(), o_rew) and np.array_equal(done.cpu().numpy(), o_done)
    assert float(big[:off].min()) == -7.0 and float(big[off + E * D:].max()) == -7.0 and float(big[off + E * D:].min()) == -7.0
    assert bits_equal(g.get_state_blob(), o.blob)
    g.close()


def test_original_env_variant():
    """CN_FLAG_ENV_ORIGINAL (environment_stage_1_original.py: 363-wide row, goal-relative heading / distance, its own
    reward): same simulator and state, different observation writer and reward -- bit-exact against the oracle, whose
    variant is pinned to the reference's own code by tests/golden/trace_original*.npz."""
    cfg = make_config(n_envs=300, auto_reset=True, layout_jitter=0.05, max_steps=120, env_original=True)
    assert cfg.obs_dim == 363
    n_done = _rollout(cfg, 200, seed=61)
    assert n_done > 0
    # the 20-pedestrian test room, several tiles, ragged tail
    big = test_world_20(n_envs=1030, auto_reset=True, layout_jitter=0.05)
    big.flags |= 4
    big.k_obstacles = 0
    big.heading_off_x = big.heading_off_y = 0.0
    big.collision_range = 0.105
    assert big.obs_dim == 363
    _rollout(big, 40, seed=62, check_every=10)


def test_original_env_needs_the_default_kernel(monkeypatch):
    from crowdnav_b200._lib import CrowdNavError
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    monkeypatch.setenv("CN_KERNEL", "warp")
    with pytest.raises(CrowdNavError):
        CrowdNavVecEnv(make_config(n_envs=4, env_original=True), device=0)


def test_realworld_370_layout_single_max_cp_slot():
    """environment_stage_1_nobonus_realworld.py's row: 359 + 7 + one slot with the highest-CP object (K = 1,
    CN_FLAG_TOPK_HIGHEST).  Bit-exact against the oracle; the slot must never hold a lower-CP object than K = N would."""
    from crowdnav_b200.config import realworld_layout_config
    cfg = realworld_layout_config(n_envs=256, auto_reset=True, layout_jitter=0.05, max_steps=150)
    assert cfg.obs_dim == 370 and cfg.k_obstacles == 1 and cfg.flags & 2
    n_done = _rollout(cfg, 150, seed=71)
    assert n_done > 0
