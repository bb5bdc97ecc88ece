"""CPU tests of the `risk_faithful` block (CN_FLAG_RISK_FAITHFUL).

1. The oracle restatement (oracle/cn_oracle_faithful.c) against the REFERENCE: the committed reference-in-the-loop
   traces hold, per step, the odometry and raw scan the reference's Env saw and the state row it returned; the K
   block of that row (ENV:862-907: the output of its own gradient typing, segmentation, uuid tracker and collision
   cone), its safety counters and the size of its tracker dict (tests/golden/faithful_counters.npz, written by
   tests/gen_golden_faithful.py) must be reproduced EXACTLY (K block to float64 round-off: 1e-9).
2. The float64 primitives of cn_math64.h against libm / CPython.
3. The DEVICE code of the block (crowdnav_b200/csrc/cn_faithful.h), built for the host with one lane, against the
   oracle bit for bit -- what the GPU test repeats with 32 lanes on the B200.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from crowdnav_b200.config import CN_FLAG_RISK_FAITHFUL, baseline_config, make_config
from oracle.oracle import OracleEnv
from trace_configs import trace_config

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"      # the distro compiler, like oracle/Makefile


class CnfParams(C.Structure):      # crowdnav_b200/csrc/cn_faithful_state.h
    _fields_ = [("n_rays", C.c_int32), ("k_obstacles", C.c_int32), ("topk_highest", C.c_int32), ("pad_", C.c_int32),
                ("inc_deg", C.c_double), ("max_range", C.c_double), ("min_range", C.c_double), ("dt", C.c_double),
                ("cp_radius", C.c_double), ("track_half", C.c_double)]


def _params(cfg, inc_deg=None):
    return CnfParams(cfg.n_samples - 1, cfg.k_obstacles, 1 if cfg.flags & 2 else 0, 0,
                     float(cfg.hit_angle_inc_deg) if inc_deg is None else inc_deg,
                     round(float(cfg.max_range), 6), round(float(cfg.collision_range), 6), round(float(cfg.dt), 6),
                     round(float(cfg.cp_radius), 6), 0.0505)


@pytest.mark.parametrize("name", ["c1", "train", "goal", "test20"])
def test_oracle_matches_reference_traces(oracle_lib, name):
    L = oracle_lib
    L.orf_observe.argtypes = [C.POINTER(CnfParams), C.c_void_p] + [C.c_double] * 3 + [C.c_void_p, C.c_int, C.c_void_p]
    cfg, _, _ = trace_config(name)
    z = np.load(os.path.join(GOLD, "trace_%s.npz" % name))
    cnt = np.load(os.path.join(GOLD, "faithful_counters.npz"))[name]
    NR, K = cfg.n_samples - 1, cfg.k_obstacles
    # the traces were recorded under Python 3, where UTL:113 `max_angle / (resolution - 1)` is a true division
    p = _params(cfg, inc_deg=360.0 / NR)
    trk = np.zeros(L.orf_world_words(), dtype=np.uint32)
    step, n_obj_rows = 0, 0
    for t in range(len(z["odom"])):
        step = 0 if z["episode_start"][t] > 0 else step + 1
        raw = z["scan"][t].astype(np.float64)                    # Gazebo order, +inf = no return
        clean = np.where(np.isinf(raw) | (raw > 0.6) | (raw == 0.0), 0.6, raw)[::-1][:-1].copy()   # UTL:375-392
        x, y, yaw = (float(q) for q in z["odom"][t][:3])
        kb = np.zeros(4 * K)
        L.orf_observe(C.byref(p), trk.ctypes.data, x, y, yaw, clean.ctypes.data, step, kb.ctypes.data)
        ref = z["ref_state"][t][NR + 7:]
        assert np.allclose(kb, ref, rtol=0.0, atol=1e-9), "row %d (step %d): K block\n%s\n%s" % (t, step, kb, ref)
        assert (int(trk[4]), int(trk[5]), int(trk[6]), int(trk[0])) == tuple(int(v) for v in cnt[t]), \
            "row %d: counters / tracker size %s vs reference %s" % (t, trk[[4, 5, 6, 0]], cnt[t])
        n_obj_rows += int(np.abs(ref.reshape(K, 4)[:, 2:]).max() > 0)
    assert int(trk[7]) == 0                                      # no capacity overflow
    if name != "c1":
        assert n_obj_rows > 200                                  # the traces do exercise the tracker


def test_sincos64_and_round3(oracle_lib):
    L = oracle_lib
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.uniform(-20.0, 20.0, 200000), np.linspace(-7, 7, 20001)])
    s, c = np.zeros_like(a), np.zeros_like(a)
    L.orf_sincos64(a.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), len(a))
    assert np.abs(s - np.sin(a)).max() <= 3e-16 and np.abs(c - np.cos(a)).max() <= 3e-16
    # round(x, 3): CPython rounds the exact binary value (ties cannot be produced by random doubles)
    x = np.concatenate([rng.uniform(-3.0, 3.0, 100000), rng.integers(-3000, 3000, 20000) / 1000.0 + 0.0005,
                        rng.integers(-3000, 3000, 20000) / 1000.0 + rng.uniform(-1e-12, 1e-12, 20000)])
    out = np.zeros_like(x)
    L.orf_round3(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(x))
    want = np.array([round(float(v), 3) for v in x])
    ties = np.array([abs(float(v) * 1000.0 % 1.0 - 0.5) == 0.0 and math.fmod(float(v) * 16.0, 1.0) == 0.0 for v in x])
    assert np.array_equal(out[~ties], want[~ties])


def test_milli64_is_the_ieee_division(oracle_lib):
    """cn_milli64 (k / 1000 by one multiplication and two fmas) == the IEEE division for every |k| <= 2^25."""
    oracle_lib.orf_milli_mismatches.restype = C.c_long
    oracle_lib.orf_milli_mismatches.argtypes = [C.c_long, C.c_long]
    assert oracle_lib.orf_milli_mismatches(-(1 << 25) - 1000, (1 << 25) + 1000) == 0


def _host_device_lib(tmp_path_factory):
    """g++ build of the device header with one lane (tests/faithful_host.cpp)."""
    out = os.path.join(str(tmp_path_factory.mktemp("faithful_host")), "libfaithful_host.so")
    cmd = [GXX, "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-mavx2", "-fPIC", "-shared",
           "-pthread", "-o", out, os.path.join(HERE, "faithful_host.cpp")]
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    H = C.CDLL(out)
    H.cnfh_observe.argtypes = [C.POINTER(CnfParams), C.c_void_p] + [C.c_double] * 3 + [C.c_void_p, C.c_float, C.c_int,
                                                                                    C.c_void_p]
    H.cnfh_observe_lanes.argtypes = H.cnfh_observe.argtypes + [C.c_int]
    return H


@pytest.fixture(scope="module")
def host_device(tmp_path_factory):
    return _host_device_lib(tmp_path_factory)


def _cases():
    c5 = baseline_config(4, n_envs=4, auto_reset=True)
    k1 = baseline_config(1, n_envs=12, auto_reset=True)
    k1.flags |= 2
    k1.k_obstacles = 1
    return [("c1", baseline_config(0, n_envs=6, auto_reset=True), 120, False),
            ("c2-shape", baseline_config(1, n_envs=16, auto_reset=True), 90, True),
            ("train", make_config(n_envs=8, auto_reset=True, layout_jitter=0.05, max_steps=80), 100, True),
            ("c5-shape", c5, 30, True), ("k1-highest", k1, 60, True)]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_device_code_on_host_equals_oracle(host_device, case):
    _, cfg, T, fast = case
    cfg = cfg.copy()
    cfg.flags |= CN_FLAG_RISK_FAITHFUL
    o = OracleEnv(cfg, debug=True)
    E, NR, K = cfg.n_envs, cfg.n_samples - 1, cfg.k_obstacles
    p = _params(cfg)
    trk = np.zeros((E, 396), dtype=np.uint32)
    rng = np.random.default_rng(3)

    def check(tag):
        rw = o.robot_words()
        for e in range(E):
            r = rw[e]
            x = float(np.float32(np.int32(r[0])) * np.float32(2.0 ** -24))
            y = float(np.float32(np.int32(r[1])) * np.float32(2.0 ** -24))
            yaw = float(np.float32(np.int32(r[2])) * np.float32(1.4629180792671596e-09))
            kb = np.zeros(4 * K, dtype=np.float32)
            sc = np.ascontiguousarray(o.ranges[e])
            host_device.cnfh_observe(C.byref(p), trk[e].ctypes.data, x, y, yaw, sc.ctypes.data,
                                     C.c_float(cfg.max_range), int(r[11]), kb.ctypes.data)
            assert np.array_equal(kb.view(np.uint32), o.obs[e, NR + 7:].view(np.uint32)), \
                "%s world %d: K block\n%s\n%s" % (tag, e, kb, o.obs[e, NR + 7:])
            assert np.array_equal(trk[e], o.trk()[e]), "%s world %d: tracker record" % (tag, e)

    o.reset()
    check("reset")
    for t in range(T):
        a = np.stack([rng.uniform(0, 0.22, E), rng.uniform(-2, 2, E)], 1).astype(np.float32)
        if fast:
            a[:, 0] = 0.22
            a[:, 1] = rng.uniform(-0.5, 0.5, E)
        o.step(a)
        check("step %d" % (t + 1))
    assert o.trk()[:, 0].max() > 0 or cfg.n_samples < 100       # objects were tracked
    assert int(o.trk()[:, 7].sum()) == 0


def _record_worlds(cfg, T, host_device=None, nl=64):
    """Step the oracle; per world and step yield what cn_faithful_kernel would read and must produce.  With
    `host_device`, also run the device code with `nl` lanes as `nl` concurrent threads and compare."""
    cfg = cfg.copy()
    cfg.flags |= CN_FLAG_RISK_FAITHFUL
    o = OracleEnv(cfg, debug=True)
    E, NR, K = cfg.n_envs, cfg.n_samples - 1, cfg.k_obstacles
    p = _params(cfg)
    trk = np.zeros((E, 396), dtype=np.uint32)
    rng = np.random.default_rng(3)
    recs = []

    def visit():
        rw = o.robot_words()
        for e in range(E):
            r = rw[e]
            x = float(np.float32(np.int32(r[0])) * np.float32(2.0 ** -24))
            y = float(np.float32(np.int32(r[1])) * np.float32(2.0 ** -24))
            yaw = float(np.float32(np.int32(r[2])) * np.float32(1.4629180792671596e-09))
            sc = np.ascontiguousarray(o.ranges[e])
            recs.append((trk[e].copy(), x, y, yaw, sc.copy(), int(r[11]), o.obs[e, NR + 7:].copy(), o.trk()[e].copy()))
            if host_device is not None:
                kb = np.zeros(4 * K, dtype=np.float32)
                host_device.cnfh_observe_lanes(C.byref(p), trk[e].ctypes.data, x, y, yaw, sc.ctypes.data,
                                               C.c_float(cfg.max_range), int(r[11]), kb.ctypes.data, nl)
                assert np.array_equal(kb.view(np.uint32), o.obs[e, NR + 7:].view(np.uint32))
                assert np.array_equal(trk[e], o.trk()[e])
            else:
                trk[e] = o.trk()[e]

    o.reset()
    visit()
    for t in range(T):
        o.step(np.stack([np.full(E, 0.22), rng.uniform(-0.5, 0.5, E)], 1).astype(np.float32))
        visit()
    return p, recs


@pytest.mark.parametrize("nl", [64, 32, 7])
def test_device_code_with_concurrent_lanes(host_device, nl):
    """The lanes of the warp as real threads, CNF_SYNC() as a barrier over them: same bits as the oracle."""
    _record_worlds(baseline_config(1, n_envs=6, auto_reset=True), 40, host_device, nl)
    _record_worlds(baseline_config(4, n_envs=2, auto_reset=True), 12, host_device, nl)


def test_device_code_is_race_free_under_thread_sanitizer(tmp_path):
    """64 lanes as 64 threads under -fsanitize=thread: every access to the world's scratch must be ordered by a
    CNF_SYNC().  (Checked once by hand that removing one sync is reported.)"""
    exe = str(tmp_path / "faithful_tsan")
    cmd = [GXX, "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-mavx2", "-fsanitize=thread",
           "-pthread", "-o", exe, os.path.join(HERE, "faithful_host_main.cpp")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime here: " + res.stdout[-200:])
    cfg = baseline_config(1, n_envs=6, auto_reset=True)
    p, recs = _record_worlds(cfg, 40)
    path = str(tmp_path / "records.bin")
    with open(path, "wb") as f:
        f.write(np.array([len(recs), cfg.n_samples - 1, cfg.k_obstacles], dtype=np.int32).tobytes())
        f.write(bytes(p))
        for trk_in, x, y, yaw, sc, step, kb, trk_out in recs:
            f.write(trk_in.tobytes())
            f.write(np.array([x, y, yaw], dtype=np.float64).tobytes())
            f.write(sc.astype(np.float32).tobytes())
            f.write(np.array([step], dtype=np.int32).tobytes())
            f.write(kb.astype(np.float32).tobytes())
            f.write(trk_out.tobytes())
    res = subprocess.run([exe, path], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if "FATAL: ThreadSanitizer" in res.stdout or "records," not in res.stdout:
        pytest.skip("ThreadSanitizer cannot run here: " + res.stdout[-200:])   # e.g. an address-space layout it rejects
    assert "ThreadSanitizer: data race" not in res.stdout, res.stdout[-3000:]
    assert res.returncode == 0, res.stdout[-2000:]


def _synth_scan(rng, n, kind):
    """Synthetic cleaned scans the simulator rarely produces: blobs, walls with objects in front, salt-and-pepper
    returns, and (kind 3) more small objects than the tracker / confirmed-list capacities."""
    s = np.full(n, 0.6, np.float32)
    if kind == 0:
        for _ in range(rng.integers(0, 12)):
            c, w, d = rng.integers(0, n), rng.integers(1, 40), rng.uniform(0.09, 0.59)
            idx = (c + np.arange(w)) % n
            s[idx] = np.minimum(s[idx], (d + rng.normal(0, 0.002, w) + 0.05 * ((np.arange(w) - w / 2) / w) ** 2).astype(np.float32))
    elif kind == 1:
        for _ in range(rng.integers(1, 3)):
            t0, d, th = rng.uniform(0, 2 * np.pi), rng.uniform(0.1, 0.55), np.arange(n) * (2 * np.pi / n)
            c = np.cos(th - t0)
            s = np.minimum(s, np.where(c > 0.05, d / np.maximum(c, 1e-3), 9.0).astype(np.float32))
        for _ in range(rng.integers(0, 6)):
            c, w, d = rng.integers(0, n), rng.integers(3, 25), rng.uniform(0.09, 0.5)
            idx = (c + np.arange(w)) % n
            s[idx] = np.minimum(s[idx], np.float32(d))
    elif kind == 2:
        m = rng.uniform(size=n) < rng.uniform(0.1, 0.9)
        s[m] = rng.uniform(0.08, 0.6, m.sum()).astype(np.float32)
        if rng.uniform() < 0.5:
            s = np.round(s, 2)
    else:
        i = 0
        while i < n - 6:
            w = rng.integers(4, 7)
            s[i:i + w] = np.float32(rng.uniform(0.15, 0.55)) + rng.normal(0, 0.0005, w).astype(np.float32)
            i += w + rng.integers(1, 4)
    return np.clip(s, np.float32(0.08), np.float32(0.6)).astype(np.float32)


@pytest.mark.parametrize("n,K,nl", [(359, 8, 1), (36, 3, 1), (720, 16, 1), (1024, 64, 1), (359, 8, 64)])
def test_device_code_fuzz_on_synthetic_scans(oracle_lib, host_device, n, K, nl):
    """Oracle vs device code (host build) on scan sequences far outside what the simulator produces, including more
    objects than the capacities (the overflow handling must agree too): K block and tracker record bit for bit."""
    L = oracle_lib
    L.orf_observe.argtypes = [C.POINTER(CnfParams), C.c_void_p] + [C.c_double] * 3 + [C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(n + K + nl)
    p = CnfParams(n, K, int(K == 3), 0, 1.0 if n == 359 else 360.0 / n, 0.6, 0.12, 0.15, 0.178, 0.0505)
    overflow = tracked = 0
    for ep in range(8 if nl == 1 else 2):
        ta, tb = np.zeros(396, np.uint32), np.zeros(396, np.uint32)
        x, y, yaw = rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-3.1, 3.1)
        kind = ep % 4
        base = _synth_scan(rng, n, kind)
        for t in range(30 if nl == 1 else 10):
            base = _synth_scan(rng, n, kind) if (kind == 3 or rng.uniform() < 0.3) else np.roll(base, rng.integers(-2, 3))
            sc = np.clip(base + rng.normal(0, 0.001, n).astype(np.float32) * (base < 0.6), np.float32(0.08), np.float32(0.6))
            sc = sc.astype(np.float32)
            s64 = np.where(sc >= np.float32(0.6), 0.6, sc.astype(np.float64))
            ka, kb = np.zeros(4 * K), np.zeros(4 * K, np.float32)
            L.orf_observe(C.byref(p), ta.ctypes.data, x, y, yaw, s64.ctypes.data, t, ka.ctypes.data)
            host_device.cnfh_observe_lanes(C.byref(p), tb.ctypes.data, x, y, yaw, sc.ctypes.data, C.c_float(0.6), t,
                                           kb.ctypes.data, nl)
            assert np.array_equal(ka.astype(np.float32).view(np.uint32), kb.view(np.uint32)), (ep, t, ka, kb)
            assert np.array_equal(ta, tb), (ep, t, np.nonzero(ta != tb)[0][:10])
            v = rng.uniform(0, 0.22)
            x = float(np.float32(x + v * 0.15 * np.cos(yaw)))
            y = float(np.float32(y + v * 0.15 * np.sin(yaw)))
            yaw = float(np.float32(yaw + rng.uniform(-0.3, 0.3)))
            tracked = max(tracked, int(ta[0]))
        overflow += int(ta[7])
    assert tracked > 2
    if n >= 359 and nl == 1:
        assert overflow > 0 and tracked == 32                    # the capacity paths were exercised


def test_faithful_env_properties():
    """The flag only replaces the K block and the counters: every other column, reward, done and the three base
    planes of the state are those of the default (`risk_intended`) environment."""
    base = baseline_config(1, n_envs=8, auto_reset=True)
    fa = base.copy()
    fa.flags |= CN_FLAG_RISK_FAITHFUL
    a_env, b_env = OracleEnv(base), OracleEnv(fa)
    a_env.reset()
    b_env.reset()
    rng = np.random.default_rng(9)
    NR, K = base.n_samples - 1, base.k_obstacles
    differs = 0
    for t in range(60):
        act = np.stack([np.full(8, 0.22), rng.uniform(-0.5, 0.5, 8)], 1).astype(np.float32)
        oa, ra, da = a_env.step(act)
        ob, rb, db = b_env.step(act)
        assert np.array_equal(oa[:, :NR + 7], ob[:, :NR + 7]) and np.array_equal(ra, rb) and np.array_equal(da, db)
        n = 16 + 8 * 16 + 2 * 8 * base.n_peds * 4
        assert np.array_equal(a_env.blob[16:n], b_env.blob[16:n])
        differs += int(not np.array_equal(oa[:, NR + 7:], ob[:, NR + 7:]))
        # every slot is either padding (robot pose, zero velocity) or a point within LiDAR range of the robot
        blk = ob[:, NR + 7:].reshape(8, K, 4)
        d = np.hypot(blk[:, :, 0] - ob[:, None, NR + 2], blk[:, :, 1] - ob[:, None, NR + 3])
        assert (d <= 0.6 + 0.25).all()
    assert b_env.blob.size == a_env.blob.size + 8 * 396
    assert differs > 0
