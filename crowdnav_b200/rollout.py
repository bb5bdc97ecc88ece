"""Vectorised TD3 rollout / replay on the device.

The reference's agent acts on ONE observation per call (`state.unsqueeze(0)`,
`action[0, 0]`, TD3:196-223), keeps its replay buffer as a Python list
(TD3:19-37) and pulls every action through `.cpu().numpy()` (TD3:206).  At
10^8 env-steps/s that is the bottleneck, so this module restates the same agent
batch-wise, everything resident on the GPU:

  * `TD3Actor` / `TD3Critic`  -- the reference networks (TD3:81-126): 3 x Linear,
    hidden 256, sigmoid * 0.22 on the linear velocity, tanh * 2.0 on the angular one;
    `load_reference_actor` reads the reference's `torch.save(state_dict)` checkpoints
    (or the npz fixture made from one) unchanged.
  * `explore`                 -- Gaussian exploration sigma = 1, clipped to the action box (TD3:67-78, 209-215).
  * `ReplayRing`              -- tensor ring buffer; rows whose step was an auto-reset (done == 2) are dropped.
  * `TD3Learner`              -- twin critics, target smoothing, delayed policy update (TD3:225-285).
  * `collect`                 -- batched rollout loop over CrowdNavVecEnv.

PyTorch only; the env step underneath is the CUDA library.
"""
from __future__ import annotations

import copy
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

MAX_LIN_VEL = 0.22   # TD3DRV:67
MAX_ANG_VEL = 2.0    # TD3DRV:68


class TD3Actor(nn.Module):
    """TD3:81-106."""

    def __init__(self, num_inputs: int, num_actions: int = 2, hidden_size: int = 256,
                 max_lin_vel: float = MAX_LIN_VEL, max_ang_vel: float = MAX_ANG_VEL):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden_size)
        self.linear2 = nn.Linear(hidden_size, hidden_size)
        self.linear3 = nn.Linear(hidden_size, num_actions)
        self.max_lin_vel, self.max_ang_vel = max_lin_vel, max_ang_vel

    def forward(self, state: torch.Tensor) -> torch.Tensor:
        x = F.relu(self.linear1(state))
        x = F.relu(self.linear2(x))
        a = self.linear3(x)
        return torch.stack([torch.sigmoid(a[:, 0]) * self.max_lin_vel, torch.tanh(a[:, 1]) * self.max_ang_vel], dim=1)


class TD3Critic(nn.Module):
    """TD3:109-126."""

    def __init__(self, num_inputs: int, num_actions: int = 2, hidden_size: int = 256):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs + num_actions, hidden_size)
        self.linear2 = nn.Linear(hidden_size, hidden_size)
        self.linear3 = nn.Linear(hidden_size, 1)

    def forward(self, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        x = torch.cat([state, action], 1)
        x = F.relu(self.linear1(x))
        x = F.relu(self.linear2(x))
        return self.linear3(x)


def load_reference_actor(path: str, device: torch.device | str = "cpu") -> TD3Actor:
    """Load one of the reference's actor checkpoints (models/td3/**/td3_actor_model_ep*.pt, a plain state_dict)
    or the .npz fixture converted from one (tests/gen_golden.py)."""
    if path.endswith(".npz"):
        z = np.load(path)
        sd = {k: torch.from_numpy(z[k]) for k in z.files}
    else:
        sd = torch.load(path, map_location="cpu")
    actor = TD3Actor(sd["linear1.weight"].shape[1], sd["linear3.weight"].shape[0], sd["linear1.weight"].shape[0])
    actor.load_state_dict(sd)
    return actor.to(device).eval()


def explore(action: torch.Tensor, sigma: float = 1.0, generator: torch.Generator | None = None) -> torch.Tensor:
    """GaussianExploration (TD3:67-78) + the clip of Agent.act (TD3:209-215), batched."""
    noise = torch.randn(action.shape, device=action.device, generator=generator) * sigma
    a = action + noise
    return torch.stack([a[:, 0].clamp(0.0, MAX_LIN_VEL), a[:, 1].clamp(-MAX_ANG_VEL, MAX_ANG_VEL)], dim=1).contiguous()


class ReplayRing:
    """Device-resident ring buffer replacing the Python-list ReplayBuffer (TD3:19-37)."""

    def __init__(self, capacity: int, obs_dim: int, device: torch.device):
        self.capacity, self.size, self.pos = capacity, 0, 0
        self.state = torch.empty((capacity, obs_dim), dtype=torch.float32, device=device)
        self.next_state = torch.empty((capacity, obs_dim), dtype=torch.float32, device=device)
        self.action = torch.empty((capacity, 2), dtype=torch.float32, device=device)
        self.reward = torch.empty((capacity, 1), dtype=torch.float32, device=device)
        self.done = torch.empty((capacity, 1), dtype=torch.float32, device=device)

    def add_batch(self, state, action, reward, next_state, done) -> int:
        """Append the transitions of one vectorised step; auto-reset rows (done == 2) carry no transition."""
        keep = done != 2
        n = int(keep.sum().item())
        if n == 0:
            return 0
        idx = (self.pos + torch.arange(n, device=state.device)) % self.capacity
        self.state[idx] = state[keep]
        self.next_state[idx] = next_state[keep]
        self.action[idx] = action[keep]
        self.reward[idx] = reward[keep].unsqueeze(1)
        self.done[idx] = (done[keep] == 1).float().unsqueeze(1)
        self.pos = (self.pos + n) % self.capacity
        self.size = min(self.size + n, self.capacity)
        return n

    def sample(self, batch_size: int, generator: torch.Generator | None = None):
        idx = torch.randint(0, self.size, (batch_size,), device=self.state.device, generator=generator)
        return self.state[idx], self.action[idx], self.reward[idx], self.next_state[idx], self.done[idx]

    def __len__(self) -> int:
        return self.size


class TD3Learner:
    """Agent.learn (TD3:225-285): clipped double-Q targets with target-policy smoothing, delayed actor update,
    Polyak averaging.  Hyper-parameters default to the reference's (configs/td3.yaml, TD3DRV:62-72)."""

    def __init__(self, obs_dim: int, device: torch.device, hidden: int = 256, actor_lr: float = 3e-4,
                 critic_lr: float = 3e-4, gamma: float = 0.99, tau: float = 0.005, noise_std: float = 0.2,
                 noise_clip: float = 0.5, policy_update: int = 2):
        self.actor = TD3Actor(obs_dim, 2, hidden).to(device)
        self.critic1, self.critic2 = TD3Critic(obs_dim, 2, hidden).to(device), TD3Critic(obs_dim, 2, hidden).to(device)
        self.t_actor, self.t_critic1, self.t_critic2 = (copy.deepcopy(m) for m in (self.actor, self.critic1, self.critic2))
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=actor_lr)
        self.opt_c1 = torch.optim.Adam(self.critic1.parameters(), lr=critic_lr)
        self.opt_c2 = torch.optim.Adam(self.critic2.parameters(), lr=critic_lr)
        self.gamma, self.tau, self.noise_std, self.noise_clip, self.policy_update = gamma, tau, noise_std, noise_clip, policy_update
        self.updates = 0

    def _soft(self, net, target):
        with torch.no_grad():
            for p, tp in zip(net.parameters(), target.parameters()):
                tp.mul_(1.0 - self.tau).add_(p, alpha=self.tau)

    def learn(self, batch) -> Dict[str, float]:
        state, action, reward, next_state, done = batch
        with torch.no_grad():
            na = self.t_actor(next_state)
            noise = (torch.randn_like(na) * self.noise_std).clamp(-self.noise_clip, self.noise_clip)
            na = na + noise
            na = torch.stack([na[:, 0].clamp(0.0, MAX_LIN_VEL), na[:, 1].clamp(-MAX_ANG_VEL, MAX_ANG_VEL)], 1)
            tq = torch.min(self.t_critic1(next_state, na), self.t_critic2(next_state, na))
            target = reward + (1.0 - done) * self.gamma * tq
        l1 = F.mse_loss(self.critic1(state, action), target)
        l2 = F.mse_loss(self.critic2(state, action), target)
        self.opt_c1.zero_grad(set_to_none=True); l1.backward(); self.opt_c1.step()
        self.opt_c2.zero_grad(set_to_none=True); l2.backward(); self.opt_c2.step()
        out = {"critic1": float(l1.detach()), "critic2": float(l2.detach())}
        self.updates += 1
        if self.updates % self.policy_update == 0:
            la = -self.critic1(state, self.actor(state)).mean()
            self.opt_actor.zero_grad(set_to_none=True); la.backward(); self.opt_actor.step()
            self._soft(self.actor, self.t_actor)
            self._soft(self.critic1, self.t_critic1)
            self._soft(self.critic2, self.t_critic2)
            out["actor"] = float(la.detach())
        return out


@torch.no_grad()
def collect(env, actor: nn.Module, steps: int, sigma: float = 0.0, replay: ReplayRing | None = None,
            generator: torch.Generator | None = None) -> Dict[str, float]:
    """Run `steps` vectorised env steps of `env` (a CrowdNavVecEnv with auto-reset) under `actor`.

    Returns episode statistics in the reference's terms (UTL:53-64): episodes finished, successes, mean
    undiscounted return per finished episode, mean length."""
    E, dev = env.E, env.device
    obs = env.obs.clone()
    ret = torch.zeros(E, device=dev)
    length = torch.zeros(E, device=dev)
    n_eps = n_succ = 0
    sum_ret = sum_len = 0.0
    for _ in range(steps):
        a = actor(obs)
        a = explore(a, sigma, generator) if sigma > 0 else a.contiguous()
        nobs, r, d = env.step(a)
        if replay is not None:
            replay.add_batch(obs, a, r, nobs, d)
        live = d != 2
        ret += torch.where(live, r, torch.zeros_like(r))
        length += live.float()
        ended = d == 1
        if ended.any():
            succ = env.counters()[:, 0] == 1
            n_eps += int(ended.sum())
            n_succ += int((ended & succ).sum())
            sum_ret += float(ret[ended].sum())
            sum_len += float(length[ended].sum())
            ret[ended] = 0.0
            length[ended] = 0.0
        obs = nobs.clone()
    return {"episodes": n_eps, "successes": n_succ, "success_rate": n_succ / max(n_eps, 1),
            "mean_return": sum_ret / max(n_eps, 1), "mean_length": sum_len / max(n_eps, 1)}
