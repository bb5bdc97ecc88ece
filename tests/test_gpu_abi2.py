"""ABI v2 on the B200, through ctypes: cn_step_n, library-owned step graphs, the host-buffer variants, blob header
checks, and two handles on two devices in ONE process (device guards + per-device shared-memory attributes)."""
import numpy as np
import pytest

from crowdnav_b200.config import baseline_config
from parity_util import bits_equal, random_actions

pytestmark = pytest.mark.gpu


def _env(cfg, device=0):
    from crowdnav_b200.vec_env import CrowdNavVecEnv
    return CrowdNavVecEnv(cfg, device=device)


def test_step_n_and_graph_equal_single_steps():
    import torch
    cfg = baseline_config(1, n_envs=500, auto_reset=True)
    a, b, c = _env(cfg), _env(cfg), _env(cfg)
    rng = np.random.default_rng(3)
    acts = torch.from_numpy(np.stack([random_actions(rng, 500) for _ in range(12)])).cuda()
    for e in (a, b, c):
        e.reset()
    rews, dones = [], []
    for i in range(12):
        _, r, d = a.step(acts[i])
        rews.append(r.clone()); dones.append(d.clone())
    _, rn, dn = b.step_n(acts)
    g = c.make_graph(acts)
    _, rg, dg = g.launch()
    torch.cuda.synchronize()
    assert torch.equal(rn, torch.stack(rews)) and torch.equal(dn, torch.stack(dones))
    assert torch.equal(rg, rn) and torch.equal(dg, dn)
    assert torch.equal(a.obs, b.obs) and torch.equal(a.obs, c.obs)
    assert bits_equal(a.get_state_blob(), b.get_state_blob()) and bits_equal(a.get_state_blob(), c.get_state_blob())
    assert a.launch_count == b.launch_count == c.launch_count == 13
    g.launch()                                            # a second replay goes on from the new state
    for i in range(12):
        a.step(acts[i])
    torch.cuda.synchronize()
    assert bits_equal(a.get_state_blob(), c.get_state_blob()) and c.launch_count == 25
    # action repeat: the same batch for 5 control periods
    rep = torch.empty((5, 500), dtype=torch.float32, device="cuda")
    a.set_state_blob(b.get_state_blob())
    b.step_n(acts[0], rep)
    for i in range(5):
        a.step(acts[0])
    torch.cuda.synchronize()
    assert bits_equal(a.get_state_blob(), b.get_state_blob())
    for e in (a, b, c):
        e.close()


def test_host_buffer_modes_agree():
    import torch
    cfg = baseline_config(1, n_envs=256, auto_reset=True)
    e_copy, e_map, e_pipe = _env(cfg), _env(cfg), _env(cfg)
    for e in (e_copy, e_map, e_pipe):
        e.reset()
    rng = np.random.default_rng(5)
    prev = None
    for t in range(25):
        act = random_actions(rng, 256)
        o1, r1, d1 = (x.copy() for x in e_copy.step_host(act, mode="copy"))
        o2, r2, d2 = (x.copy() for x in e_map.step_host(act, mode="mapped"))
        assert bits_equal(o1, o2) and bits_equal(r1, r2) and bits_equal(d1, d2), "mapped != copy at step %d" % t
        out = e_pipe.step_host_pipelined(act)
        if t == 0:
            assert out is None
        else:
            assert all(bits_equal(x, y) for x, y in zip(out, prev)), "pipelined results are not those of step t-1 (t=%d)" % t
        prev = (o1, r1, d1)
    last = e_pipe.flush_host_pipeline()
    assert all(bits_equal(x, y) for x, y in zip(last, prev))
    torch.cuda.synchronize()
    assert bits_equal(e_copy.get_state_blob(), e_map.get_state_blob()) and bits_equal(e_copy.get_state_blob(), e_pipe.get_state_blob())


def test_blob_header_is_checked():
    from crowdnav_b200._lib import CrowdNavError
    cfg = baseline_config(1, n_envs=64)
    e = _env(cfg)
    e.reset()
    blob = e.get_state_blob()
    e.set_state_blob(blob)
    for word, what in ((1, "layout version"), (4, "n_samples"), (5, "k_obstacles"), (6, "risk-block mode"), (7, "tracker size")):
        bad = blob.copy()
        bad[word] += 1
        with pytest.raises(CrowdNavError):
            e.set_state_blob(bad)
    other = baseline_config(1, n_envs=64)
    other.flags |= 2                                       # CN_FLAG_TOPK_HIGHEST does not change the state: accepted
    e2 = _env(other)
    e2.set_state_blob(blob)
    e.close(); e2.close()


def test_two_handles_on_two_devices_in_one_process():
    """Needs 2 GPUs (skipped otherwise).  c5-shaped worlds use > 48 KB of dynamic shared memory per CTA, so the second
    device's launch fails unless the attribute is configured per device; calls interleave without the caller ever
    selecting a device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle.oracle import OracleEnv
    cfg = baseline_config(4, n_envs=96, auto_reset=True)
    torch.cuda.set_device(0)
    e0, e1 = _env(cfg, 0), _env(cfg, 1)
    o = OracleEnv(cfg)
    assert torch.cuda.current_device() == 0               # cn_create put the caller's device back
    e0.reset(); e1.reset(); o.reset()
    rng = np.random.default_rng(11)
    for t in range(10):
        act = random_actions(rng, 96)
        e0.step(torch.from_numpy(act).to("cuda:0"))
        e1.step(torch.from_numpy(act).to("cuda:1"))
        o.step(act)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0
    assert bits_equal(e0.obs.cpu().numpy(), o.obs) and bits_equal(e1.obs.cpu().numpy(), o.obs)
    assert bits_equal(e0.get_state_blob(), o.blob) and bits_equal(e1.get_state_blob(), o.blob)
    e0.close(); e1.close()
