"""The C-ABI library loads without a GPU and exports every symbol include/crowdnav.h declares."""
import ctypes as C
import os
import re

import pytest

from crowdnav_b200 import _lib
from crowdnav_b200.config import CnConfig, make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    _lib.build_library()
    return _lib.load()


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "crowdnav.h")).read()
    declared = set(re.findall(r"\b(cn_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.ABI_SYMBOLS), "include/crowdnav.h and _lib.ABI_SYMBOLS disagree"
    for sym in sorted(declared):
        assert hasattr(L, sym), "libcrowdnav.so does not export %s" % sym


def test_struct_layout_and_pure_entry_points(L):
    cfg = CnConfig()
    assert L.cn_config_default(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(CnConfig), "ctypes mirror of cn_config is out of sync with the header"
    assert (cfg.n_peds, cfg.n_samples, cfg.k_obstacles) == (14, 360, 8)
    assert L.cn_obs_dim(C.byref(cfg)) == 398 == cfg.obs_dim
    py = make_config()
    for f in ("dt", "room_xmin", "room_xmax", "room_ymin", "room_ymax", "start_x", "start_y", "start_yaw", "goal_x",
              "goal_y", "heading_off_x", "heading_off_y", "max_range", "collision_range", "sensor_min_range",
              "sensor_sweep", "mount_x", "hit_angle_inc_deg", "ped_radius", "robot_radius", "cp_radius",
              "waypoint_radius", "goal_box"):
        assert getattr(cfg, f) == getattr(py, f), f
    assert L.cn_blob_bytes(C.byref(cfg)) == 4 * (16 + 16 + 2 * 14 * 4)
    assert L.cn_abi_version() == 2


def test_risk_faithful_flag_in_the_abi(L):
    """CN_FLAG_RISK_FAITHFUL: same row width, the blob grows by one tracker record (396 words) per world, the flag
    is refused together with the original environment (which has no K block)."""
    from crowdnav_b200.config import CN_FLAG_RISK_FAITHFUL
    hdr = open(os.path.join(ROOT, "include", "crowdnav.h")).read()
    assert int(re.search(r"#define CN_FLAG_RISK_FAITHFUL\s+(\d+)u", hdr).group(1)) == CN_FLAG_RISK_FAITHFUL == 8
    base, fa = make_config(n_envs=5), make_config(n_envs=5, risk_faithful=True)
    assert fa.flags & 8 and L.cn_obs_dim(C.byref(fa)) == L.cn_obs_dim(C.byref(base)) == 398
    assert L.cn_blob_bytes(C.byref(fa)) == L.cn_blob_bytes(C.byref(base)) + 5 * 396 * 4
    with pytest.raises(ValueError):
        make_config(env_original=True, risk_faithful=True)
    bad = make_config(env_original=True)
    bad.flags |= 8
    h = C.c_void_p()
    assert L.cn_create(C.byref(bad), 0, C.byref(h)) == -1


def test_errors_are_codes_not_crashes(L):
    h = C.c_void_p()
    bad = make_config()
    bad.struct_size = 12
    assert L.cn_create(C.byref(bad), 0, C.byref(h)) == -1 and b"struct_size" in L.cn_last_error()
    bad = make_config(n_peds=14)
    bad.n_peds = 100
    assert L.cn_create(C.byref(bad), 0, C.byref(h)) == -1
    assert L.cn_step(None, None, None, None, None, None) == -1
    assert L.cn_reset(None, None, None, None) == -1
    assert L.cn_destroy(None) == 0
    import torch
    if not torch.cuda.is_available():
        # no device here: creation must fail loudly with a CUDA error code, never fall back to a CPU path
        rc = L.cn_create(C.byref(make_config()), 0, C.byref(h))
        assert rc in (-2, -1) and L.cn_last_error()
        from crowdnav_b200.vec_env import CrowdNavVecEnv
        with pytest.raises(_lib.CrowdNavError):
            CrowdNavVecEnv(make_config())


def test_product_never_imports_the_oracle():
    """Nothing under crowdnav_b200/ may reference oracle/ (the product must not route through the checker)."""
    pkg = os.path.join(ROOT, "crowdnav_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "cn_oracle" not in txt.replace(
                    "oracle/cn_oracle.c", ""), f


def test_tile_plan_for_the_baseline_configs(L):
    """Host-side launch planning of the default (flat) kernel: the tile must fit a fifth of a B200 SM's shared
    memory (5 resident CTAs per SM), keep the row block 16-byte aligned for the bulk store, and give one wave when the
    batch allows it."""
    from crowdnav_b200.config import baseline_config
    n_sms, smem_sm = 148, 233472            # B200
    plans = {}
    for idx, name in ((1, "c2"), (2, "c3"), (3, "c4"), (4, "c5"), (0, "c1")):
        cfg = baseline_config(idx)
        if name == "c4":
            cfg.n_envs = 8192               # per-GPU shard of the 65536-world config
        tile, threads, smem = C.c_int(), C.c_int(), C.c_size_t()
        assert L.cn_plan_tile(C.byref(cfg), n_sms, smem_sm, C.byref(tile), C.byref(threads), C.byref(smem)) == 0
        plans[name] = (tile.value, threads.value, smem.value)
        assert threads.value == 256 and 1 <= tile.value <= 16
        assert smem.value <= smem_sm // 5 - 1024
        assert (tile.value * cfg.obs_dim) % 4 == 0 and tile.value % 2 == 0, "row block must be able to leave by bulk store"
        assert tile.value * cfg.n_peds <= 0x3FFF
    # c2 fits one wave of 5 CTAs/SM with 6-world tiles; c3 fills two waves best with 12-world tiles
    assert plans["c2"][0] == 6 and (4096 + 5) // 6 <= 5 * n_sms
    assert plans["c3"][0] == 12
    assert (8192 + plans["c4"][0] - 1) // plans["c4"][0] <= 2 * 5 * n_sms
    # the staging tile of the pipelined gather (CN_FLAG_GATHER_STAGE) still fits
    staged = baseline_config(1)
    staged.flags |= 16
    tile, threads, smem = C.c_int(), C.c_int(), C.c_size_t()
    assert L.cn_plan_tile(C.byref(staged), n_sms, smem_sm, C.byref(tile), C.byref(threads), C.byref(smem)) == 0
    assert smem.value <= smem_sm // 5 - 1024 and (4096 + tile.value - 1) // tile.value <= 5 * n_sms
    bad = make_config()
    bad.n_peds = 1000
    assert L.cn_plan_tile(C.byref(bad), n_sms, smem_sm, C.byref(tile), C.byref(threads), C.byref(smem)) == -1


def test_tile_plan_of_the_direct_rows_instance(L):
    """Plain steps into device memory run the direct-rows instance (no staging tile: ~2.0 instead of ~3.8 KB of shared
    memory per world).  Its planner must give BASELINE configs[2] ONE wave (that is the point of the layout: 863 CTAs of
    19 worlds on 6 x 148 slots), keep one wave for c2 / the c4 shard, use 384-thread CTAs for the 50-pedestrian config,
    and never exceed the CTA's share of a B200 SM."""
    from crowdnav_b200.config import baseline_config
    n_sms, smem_sm = 148, 233472
    plans = {}
    for idx, name in ((1, "c2"), (2, "c3"), (3, "c4"), (4, "c5"), (0, "c1")):
        cfg = baseline_config(idx)
        if name == "c4":
            cfg.n_envs = 8192
        tile, threads, smem = C.c_int(), C.c_int(), C.c_size_t()
        assert L.cn_plan_tile_direct(C.byref(cfg), n_sms, smem_sm, C.byref(tile), C.byref(threads), C.byref(smem)) == 0
        plans[name] = (tile.value, threads.value, smem.value)
        ctas = {256: 6, 384: 4}[threads.value]
        assert 1 <= tile.value <= 32 and smem.value <= smem_sm // ctas - 1024
        assert ctas * (smem.value + 1024) <= smem_sm, "the resident CTAs (+ 1 KB each reserved by the driver) must fit the SM"
        staged_tile, st, ss = C.c_int(), C.c_int(), C.c_size_t()
        assert L.cn_plan_tile(C.byref(cfg), n_sms, smem_sm, C.byref(staged_tile), C.byref(st), C.byref(ss)) == 0
        if name != "c1":        # (36-ray rows are too small to matter)
            assert smem.value / tile.value < 0.62 * ss.value / staged_tile.value, "no row block in the direct layout"
    assert plans["c2"][:2] == (6, 256)
    assert plans["c3"][:2] == (19, 256) and (16384 + 18) // 19 <= 6 * n_sms          # one wave
    assert plans["c4"][1] == 256 and (8192 + plans["c4"][0] - 1) // plans["c4"][0] <= 6 * n_sms
    assert plans["c5"][1] == 384


def test_original_env_flag_row_width(L):
    """CN_FLAG_ENV_ORIGINAL: (R-1) + 4 columns, K must be 0 (library and ctypes mirror agree)."""
    cfg = make_config(env_original=True)
    assert cfg.flags & 4 and cfg.k_obstacles == 0
    assert L.cn_obs_dim(C.byref(cfg)) == 363 == cfg.obs_dim
    tile, threads, smem = C.c_int(), C.c_int(), C.c_size_t()
    assert L.cn_plan_tile(C.byref(cfg), 148, 233472, C.byref(tile), C.byref(threads), C.byref(smem)) == 0
    cfg.k_obstacles = 8
    assert L.cn_plan_tile(C.byref(cfg), 148, 233472, C.byref(tile), C.byref(threads), C.byref(smem)) == -1
