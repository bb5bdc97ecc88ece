"""Host logic of the vectorised rollout (crowdnav_b200/rollout.py) on CPU tensors: the sync-free replay ring, the SAC
actor's checkpoint layout, the TD3 target (TD3:247-250)."""
import numpy as np
import torch

from crowdnav_b200.rollout import ReplayRing, SACActor, TD3Actor, TD3Learner


def _batch(n, d, base, done):
    s = torch.arange(base, base + n, dtype=torch.float32).unsqueeze(1).repeat(1, d)
    return s, s[:, :2] + 0.25, s[:, 0] + 0.5, s + 1000.0, torch.tensor(done, dtype=torch.uint8)


def test_ring_appends_in_order_and_skips_auto_reset_rows():
    ring = ReplayRing(8, 3, torch.device("cpu"))
    ring.add_batch(*_batch(5, 3, 0, [0, 2, 1, 0, 2]))          # rows 0, 2, 3 kept
    assert len(ring) == 3 and ring.pos == 3
    assert ring.state[:3, 0].tolist() == [0.0, 2.0, 3.0] and ring.done[:3, 0].tolist() == [0.0, 1.0, 0.0]
    ring.add_batch(*_batch(7, 3, 10, [0] * 7))                 # wraps: slots 3..7, 0, 1
    assert len(ring) == 8 and ring.pos == 2
    assert ring.state[:8, 0].tolist() == [15.0, 16.0, 3.0, 10.0, 11.0, 12.0, 13.0, 14.0]
    # every field of a slot belongs to the same transition
    assert torch.equal(ring.next_state[:8], ring.state[:8] + 1000.0) and torch.equal(ring.reward[:8, 0], ring.state[:8, 0] + 0.5)
    assert torch.equal(ring.action[:8], ring.state[:8, :2] + 0.25)
    s, a, r, s2, d = ring.sample(64)
    assert s.shape == (64, 3) and torch.equal(s2, s + 1000.0)


def test_ring_batch_larger_than_capacity_keeps_the_last_rows_consistently():
    ring = ReplayRing(4, 2, torch.device("cpu"))
    done = [0, 2, 0, 0, 0, 2, 0, 0, 0, 0]                       # 8 transitions, capacity 4
    ring.add_batch(*_batch(10, 2, 0, done))
    assert len(ring) == 4
    kept = sorted(ring.state[:4, 0].tolist())
    assert kept == [6.0, 7.0, 8.0, 9.0]
    assert torch.equal(ring.next_state[:4], ring.state[:4] + 1000.0) and torch.equal(ring.action[:4], ring.state[:4, :2] + 0.25)
    ring.add_batch(*_batch(3, 2, 20, [2, 2, 2]))               # nothing to add
    assert len(ring) == 4 and sorted(ring.state[:4, 0].tolist()) == kept


def test_sac_actor_has_the_reference_checkpoint_layout_and_action_box():
    torch.manual_seed(0)
    actor = SACActor(398)
    assert sorted(actor.state_dict()) == sorted(
        ["linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias", "mean_linear.weight", "mean_linear.bias",
         "log_std_linear.weight", "log_std_linear.bias"])              # sac.py:51-61
    a = actor(torch.randn(512, 398))
    assert a.shape == (512, 2)
    # sigmoid(tanh(z)) * 0.22 in (0.059, 0.161); tanh(tanh(z)) * 2 in (-1.524, 1.524)   (sac.py:97-101, double squashing)
    assert float(a[:, 0].min()) > 0.0591 and float(a[:, 0].max()) < 0.1609 and float(a[:, 1].abs().max()) < 1.5232
    assert torch.equal(actor(torch.ones(4, 398), deterministic=True), actor(torch.ones(4, 398), deterministic=True))


def test_td3_target_action_is_not_clipped_to_the_action_box():
    """TD3:247-250 adds the clipped noise to the target action and does not clamp the sum."""
    torch.manual_seed(1)
    L = TD3Learner(6, torch.device("cpu"), hidden=16)
    seen = {}
    orig = L.t_critic1.forward

    def spy(state, action):
        seen["a"] = action.detach().clone()
        return orig(state, action)
    L.t_critic1.forward = spy
    with torch.no_grad():                                               # saturate the actor: v = 0.22, |w| = 2
        L.t_actor.linear3.bias.copy_(torch.tensor([50.0, 50.0]))
    b = (torch.randn(256, 6), torch.rand(256, 2), torch.randn(256, 1), torch.randn(256, 6), torch.zeros(256, 1))
    out = L.learn(b)
    assert float(seen["a"][:, 0].max()) > 0.22 and float(seen["a"][:, 1].max()) > 2.0
    assert float((seen["a"][:, 0] - 0.22).abs().max()) <= 0.5 + 1e-6    # noise_clip
    assert np.isfinite(float(out["critic1"]))
    assert isinstance(TD3Actor(6)(torch.zeros(2, 6)), torch.Tensor)
