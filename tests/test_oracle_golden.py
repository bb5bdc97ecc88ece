"""Pin the CPU oracle to the REFERENCE: golden vectors and whole-episode traces
produced by running the reference's own utils.py / Env code (tests/gen_golden.py,
tests/ref_harness.py) are replayed against oracle/cn_oracle.c.

Tolerances (written here, per BASELINE.json: ranges / reward within 1e-4 fp32):
  * observation columns other than the K-obstacle block: |diff| <= 1e-6 on every
    value (the reference computes in float64 and hands float32 to the agent; 1e-6
    is float32 representation error of values up to 3).  No rounding flips are
    tolerated on the committed traces.
  * reward and done: exact.
  * K-obstacle block: the reference's LiDAR segmentation / uuid-keyed tracker is
    order- and wall-clock-dependent, so the oracle's
    `risk_intended` restatement is compared statistically: slot occupancy and
    object positions.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.oracle import OracleEnv, lib
from crowdnav_b200.config import make_config
from trace_configs import TRACES, TRACES_ORIGINAL, trace_config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ctx(cfg):
    env = OracleEnv(cfg)
    return env, env._ctx, lib()


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- traces
@pytest.mark.parametrize("name", TRACES)
def test_trace_replay_matches_reference(name):
    t = np.load(os.path.join(GOLD, "trace_%s.npz" % name))
    cfg, _, _ = trace_config(name)
    o = OracleEnv(cfg, debug=True)
    R, K = cfg.n_samples, cfg.k_obstacles
    NR = R - 1
    n_rows = len(t["action"])
    occ_equal = occ_rows = 0
    pos_diffs = []
    for i in range(n_rows):
        if t["episode_start"][i] > 0:
            obs = o.reset()[0]
            rew, done = 0.0, 0
        else:
            a = t["action"][i].astype(np.float32).reshape(1, 2)
            obs_, r_, d_ = o.step(a)
            obs, rew, done = obs_[0], float(r_[0]), int(d_[0])
        # physics replays bit-identically: the raw scan handed to the reference is reproduced
        raw = np.full(R, np.inf, dtype=np.float32)
        hit = o.hit_ids[0] != 0xFF
        idx = (R - 1) - np.arange(NR)
        raw[idx[hit]] = o.ranges[0][hit]
        assert np.array_equal(raw, t["scan"][i]), "row %d: simulator no longer reproduces the recorded scan" % i
        ref = t["ref_state"][i]
        d = np.abs(obs[:NR + 7].astype(np.float64) - ref[:NR + 7])
        assert d.max() <= 1e-6, "row %d: column %d differs from the reference by %g" % (i, int(d.argmax()), d.max())
        if t["episode_start"][i] == 0:
            assert rew == t["ref_reward"][i], "row %d: reward %g vs reference %g" % (i, rew, t["ref_reward"][i])
            assert done == int(t["ref_done"][i]), "row %d: done differs" % i
            if done:
                assert int(o.counters()[0, 0]) == int(t["ref_success"][i]), "row %d: success flag differs" % i
        # K block, statistically
        x, y = ref[NR + 2], ref[NR + 3]
        rb = ref[NR + 7:].reshape(K, 4)
        ob = obs[NR + 7:].astype(np.float64).reshape(K, 4)
        r_occ = int(((np.abs(rb[:, 0] - x) > 1e-6) | (np.abs(rb[:, 1] - y) > 1e-6)).sum())
        o_occ = int(((np.abs(ob[:, 0] - x) > 1e-6) | (np.abs(ob[:, 1] - y) > 1e-6)).sum())
        occ_rows += 1
        occ_equal += int(r_occ == o_occ)
        if r_occ == 1 and o_occ == 1:
            pos_diffs.append(float(np.hypot(rb[0, 0] - ob[0, 0], rb[0, 1] - ob[0, 1])))
    if R == 360:
        # with 1-degree rays the reference's own segmentation sees the same objects as ideal association
        assert occ_equal / occ_rows >= 0.90, "K-block occupancy agrees on only %d/%d rows" % (occ_equal, occ_rows)
        assert len(pos_diffs) > 50 and np.median(pos_diffs) < 0.01, "object points differ: median %g" % np.median(pos_diffs)


@pytest.mark.parametrize("name", TRACES_ORIGINAL)
def test_original_env_trace_replay_matches_reference(name):
    """CN_FLAG_ENV_ORIGINAL against the reference's environment_stage_1_original.Env in the loop: every one of the
    363 columns to 1e-6 (float32 representation of the float64 row), reward / done / success exactly -- including the
    reference's reading of the row's last two entries (x, y) as "heading" and "distance" in compute_reward."""
    t = np.load(os.path.join(GOLD, "trace_%s.npz" % name))
    cfg, _, _ = trace_config(name)
    assert cfg.obs_dim == 363 == t["ref_state"].shape[1]
    o = OracleEnv(cfg, debug=True)
    R = cfg.n_samples
    NR = R - 1
    rewards = set()
    for i in range(len(t["action"])):
        if t["episode_start"][i] > 0:
            obs = o.reset()[0]
            rew, done = 0.0, 0
        else:
            obs_, r_, d_ = o.step(t["action"][i].astype(np.float32).reshape(1, 2))
            obs, rew, done = obs_[0], float(r_[0]), int(d_[0])
        raw = np.full(R, np.inf, dtype=np.float32)
        hit = o.hit_ids[0] != 0xFF
        idx = (R - 1) - np.arange(NR)
        raw[idx[hit]] = o.ranges[0][hit]
        assert np.array_equal(raw, t["scan"][i]), "row %d: simulator no longer reproduces the recorded scan" % i
        d = np.abs(obs.astype(np.float64) - t["ref_state"][i])
        assert d.max() <= 1e-6, "row %d: column %d differs from the reference by %g" % (i, int(d.argmax()), d.max())
        if t["episode_start"][i] == 0:
            assert rew == t["ref_reward"][i], "row %d: reward %g vs reference %g" % (i, rew, t["ref_reward"][i])
            assert done == int(t["ref_done"][i]), "row %d: done differs" % i
            if done:
                assert int(o.counters()[0, 0]) == int(t["ref_success"][i]), "row %d: success flag differs" % i
            rewards.add(rew)
    assert rewards & {0.0, 1.0, 2.0}, "no progress rewards in the trace"


def test_original_env_traces_cover_its_reward_terms():
    seen = set()
    for name in TRACES_ORIGINAL:
        seen |= set(np.load(os.path.join(GOLD, "trace_%s.npz" % name))["ref_reward"].tolist())
    for r in (0.0, 1.0, 2.0):
        assert r in seen, "no golden row with reward %g" % r
    assert seen & {200.0, 201.0, 202.0}, "no successful episode in the original-env traces"
    assert seen & {-200.0, -199.0, -198.0}, "no failed episode in the original-env traces"


def test_traces_cover_all_reward_terms():
    """The committed traces exercise step / distance / heading / waypoint / goal / collision terms."""
    seen = set()
    for name in TRACES:
        seen |= set(np.load(os.path.join(GOLD, "trace_%s.npz" % name))["ref_reward"].tolist())
    for r in (-2.0, -1.0, 0.0, 198.0, 199.0, 200.0, -202.0, -201.0):
        assert r in seen, "no golden row with reward %g" % r


# ----------------------------------------------------------------- utils.py vectors
def test_get_scan_ranges_golden():
    """utils.get_scan_ranges (UTL:375-392): inf / NaN / 0 / > max, reverse, drop."""
    g = np.load(os.path.join(GOLD, "utils_vectors.npz"))
    env, ctx, L = _ctx(make_config(n_samples=360))
    for raw, want in zip(g["raw"], g["cleaned"]):
        raw32 = raw.astype(np.float32)
        out = np.zeros(359, dtype=np.float32)
        L.orc_clean_scan(C.c_void_p(ctx), _fp(raw32), _fp(out))
        assert np.array_equal(out, want.astype(np.float32))


@pytest.mark.parametrize("R,key_s,key_c", [(360, "scans", "coords"), (37, "scans37", "coords37")])
def test_convert_laserscan_to_coordinate_golden(R, key_s, key_c):
    """utils.convert_laserscan_to_coordinate (UTL:110-126) incl. the Python-2 integer-degree increment."""
    g = np.load(os.path.join(GOLD, "utils_vectors.npz"))
    env, ctx, L = _ctx(make_config(n_samples=R))
    L.orc_hit_points.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]
    n_flip = n_tot = 0
    for scans, pose, want in zip(g[key_s], g["poses"], g[key_c]):
        # the oracle's yaw is a binary angle: quantise the golden yaw the same way before comparing
        th = np.uint32(int(round(pose[2] / (2 * np.pi) * 2 ** 32)) % 2 ** 32)
        s32 = scans.astype(np.float32)
        out = np.zeros((R - 1, 2), dtype=np.float32)
        L.orc_hit_points(C.c_void_p(ctx), np.float32(pose[0]), np.float32(pose[1]), th, _fp(s32), _fp(out))
        d = np.abs(out.astype(np.float64) - want)
        assert d.max() <= 1.0e-3 + 1e-6          # values are rounded to 3 dp: at most one unit in the last place
        n_flip += int((d > 1e-6).sum())
        n_tot += d.size
    assert n_flip <= 0.002 * n_tot, "%d of %d coordinates flipped a rounding boundary" % (n_flip, n_tot)


def test_collision_probability_formulas_golden():
    """utils.compute_collision_prob / compute_general_collision_prob (UTL:317-345)."""
    g = np.load(os.path.join(GOLD, "utils_vectors.npz"))
    env, ctx, L = _ctx(make_config())
    L.orc_cp_ttc.restype = C.c_float
    L.orc_cp_ttc.argtypes = [C.c_int, C.c_float, C.c_float]
    L.orc_cp_dto.restype = C.c_float
    L.orc_cp_dto.argtypes = [C.c_void_p, C.c_float]
    for ttc, want in zip(g["ttc"], g["cp_ttc"][:-1]):
        # ttc = dtc / resultant; the oracle takes (dtc > 0, resultant)
        got = L.orc_cp_ttc(1, np.float32(abs(ttc)), np.float32(1.0 if ttc > 0 else -1.0))
        assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (ttc, got, want)
    assert L.orc_cp_ttc(0, 0.0, 0.0) == g["cp_ttc"][-1] == 0.0          # None -> 0
    for d, want in zip(g["d"], g["cp_dto"]):
        got = L.orc_cp_dto(C.c_void_p(ctx), np.float32(d))
        assert abs(got - want) <= 2e-6, (d, got, want)
    # known answers quoted in SURVEY.md 8(c)
    assert abs(L.orc_cp_dto(C.c_void_p(ctx), 0.12) - 1.0) < 1e-6
    assert abs(L.orc_cp_dto(C.c_void_p(ctx), 0.36) - 0.5) < 1e-6
    assert L.orc_cp_dto(C.c_void_p(ctx), 0.6) == 0.0
    assert abs(L.orc_cp_ttc(1, 0.15, 1.0) - 1.0) < 1e-6 and abs(L.orc_cp_ttc(1, 1.5, 1.0) - 0.1) < 1e-6


def test_reference_known_answers_in_golden():
    """Constants the reference itself documents."""
    g = np.load(os.path.join(GOLD, "utils_vectors.npz"))
    # environment_stage_1_nobonus_realworld.py:103 quotes 0.0105090183944 for this quantity
    assert abs(float(g["bbox"][0]) - 0.0105090183944) < 2e-6
    d, n = g["d"], g["nscan"]
    assert n[np.argmin(np.abs(d - 0.6))] == 3                        # UTL:396-399 "3 scans at max range"


# --------------------------------------------------------------------------- waypoint
def test_local_goal_waypoints_golden():
    """utils.get_local_goal_waypoints (UTL:296-314): 64-gon ring crossing + (-gx, gy) fallback."""
    g = np.load(os.path.join(GOLD, "waypoints.npz"))
    env = OracleEnv(make_config())
    n_fallback = 0
    for a, want in zip(g["agent"], g["waypoint"]):
        wx, wy = env.waypoint(np.float32(a[0]), np.float32(a[1]))
        if want[0] == 1.0 and want[1] == 1.0:       # fallback (-(-1), 1)
            n_fallback += 1
            assert (wx, wy) == (1.0, 1.0)
        else:
            assert abs(wx - want[0]) <= 2e-6 and abs(wy - want[1]) <= 2e-6, (a, (wx, wy), want)
    assert n_fallback >= 15


# ------------------------------------------------------------------- Env methods
def test_heading_distance_boxes_golden():
    """Env.get_heading_to_goal / get_distance_to_goal / is_in_true_desired_position (ENV:191-237, 1303-1319)."""
    g = np.load(os.path.join(GOLD, "env_methods.npz"))
    env, ctx, L = _ctx(make_config())
    L.orc_distance.restype = C.c_float
    L.orc_distance.argtypes = [C.c_float] * 4
    L.orc_in_goal_box.argtypes = [C.c_void_p, C.c_float, C.c_float]
    for p, w, h, d, ig in zip(g["pose"], g["wp"], g["head"], g["dist"], g["in_goal"]):
        got = env.heading(np.float32(p[0]), np.float32(p[1]), np.float32(p[2]), np.float32(w[0]), np.float32(w[1]))
        dh = abs(got - h)
        dh = min(dh, abs(dh - 2 * np.pi))          # the single wrap of ENV:231-235 can land on either side of +-pi
        assert dh <= 2e-6, (p, w, got, h)
        assert abs(L.orc_distance(np.float32(p[0]), np.float32(p[1]), np.float32(w[0]), np.float32(w[1])) - d) <= 1e-6
        assert bool(L.orc_in_goal_box(C.c_void_p(ctx), np.float32(p[0]), np.float32(p[1]))) == bool(ig)
    # half-open edges of the goal box: (lo, hi]
    for p, want in zip(g["edge_pts"][4:], g["edge_goal"][4:]):
        assert bool(L.orc_in_goal_box(C.c_void_p(ctx), np.float32(p[0]), np.float32(p[1]))) == bool(want)


def test_reward_truth_table_golden():
    """Env.compute_reward (ENV:1046-1162): all sign combinations of heading / distance change."""
    g = np.load(os.path.join(GOLD, "env_methods.npz"))
    L = lib()
    L.orc_shaping_reward.argtypes = [C.c_float] * 4
    for ph, ch, pd, cd, done, at_goal, r, d, succ, fail in g["reward_table"]:
        shaping = L.orc_shaping_reward(np.float32(ch), np.float32(cd), np.float32(ph), np.float32(pd))
        terminal = 0 if not done else (200 if at_goal else -200)          # timeout / collision both -200
        assert shaping + terminal == r, (ph, ch, pd, cd, done, at_goal, r)
        assert shaping in (-2, -1, 0)
        assert bool(succ) == bool(done and at_goal) and bool(fail) == bool(done and not at_goal)


def test_observation_width_matches_shipped_actors():
    """First-layer widths of the shipped TD3 actors: 370/382/398/414/430 for K=1/4/8/12/16 (TD3DRV:88)."""
    for k, width in ((1, 370), (4, 382), (8, 398), (12, 414), (16, 430)):
        assert make_config(k_obstacles=k).obs_dim == width
