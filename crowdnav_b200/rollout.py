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
  * `SACActor`                -- the reference's SAC policy network (sac.py:43-106), same checkpoint layout.
  * `ReplayRing`              -- tensor ring buffer; rows whose step was an auto-reset (done == 2) are dropped;
                                 device-side cursor, no host synchronisation.
  * `TD3Learner`              -- twin critics, target smoothing, delayed policy update (TD3:225-285).
  * `collect`                 -- batched rollout loop over CrowdNavVecEnv, statistics accumulated on the device.
  * `GraphedCollector`        -- the same loop (optionally with the TD3 update) captured once as a CUDA graph.
  * `rollout_throughput`      -- env-steps/s of policy + env + replay (+ learner) for bench.py, eager or graphed.

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
    """Device-resident ring buffer replacing the Python-list ReplayBuffer (TD3:19-37).

    No host synchronisation anywhere: the write cursor and the fill level are device tensors, a vectorised step is
    appended with one scatter per field (rows that carry no transition go to a spare trash row), and `sample` draws
    its indices on the device.  `len()` is the only call that reads the fill level back."""

    def __init__(self, capacity: int, obs_dim: int, device: torch.device):
        self.capacity = capacity
        # row `capacity` is the trash row: scatter target of the rows that are dropped
        self.state = torch.zeros((capacity + 1, obs_dim), dtype=torch.float32, device=device)
        self.next_state = torch.zeros((capacity + 1, obs_dim), dtype=torch.float32, device=device)
        self.action = torch.zeros((capacity + 1, 2), dtype=torch.float32, device=device)
        self.reward = torch.zeros((capacity + 1, 1), dtype=torch.float32, device=device)
        self.done = torch.zeros((capacity + 1, 1), dtype=torch.float32, device=device)
        self._pos = torch.zeros((), dtype=torch.int64, device=device)      # next slot to write
        self._size = torch.zeros((), dtype=torch.int64, device=device)     # valid rows

    def add_batch(self, state, action, reward, next_state, done) -> None:
        """Append the transitions of one vectorised step; auto-reset rows (done == 2) carry no transition.  When one
        batch holds more transitions than the ring, only the last `capacity` of them are kept (every slot is then
        written exactly once, so the five fields of a slot always belong to the same transition)."""
        keep = done != 2
        order = torch.cumsum(keep.to(torch.int64), 0) - 1            # rank of each kept row among the kept rows
        n = order[-1] + 1
        keep = keep & (order >= n - self.capacity)
        idx = torch.where(keep, (self._pos + order) % self.capacity, torch.full_like(order, self.capacity))
        self.state[idx] = state
        self.next_state[idx] = next_state
        self.action[idx] = action
        self.reward[idx] = reward.unsqueeze(1)
        self.done[idx] = (done == 1).float().unsqueeze(1)
        self._pos.copy_((self._pos + n) % self.capacity)
        self._size.copy_(torch.clamp(self._size + n, max=self.capacity))

    def sample(self, batch_size: int, generator: torch.Generator | None = None):
        u = torch.rand((batch_size,), device=self.state.device, generator=generator)
        idx = (u * self._size.clamp(min=1).to(torch.float32)).to(torch.int64).clamp_(max=self.capacity - 1)
        return self.state[idx], self.action[idx], self.reward[idx], self.next_state[idx], self.done[idx]

    @property
    def pos(self) -> int:
        return int(self._pos.item())

    @property
    def size(self) -> int:
        return int(self._size.item())

    def __len__(self) -> int:
        return self.size


class SACActor(nn.Module):
    """The reference's SAC policy network (sac.py:43-106): shared trunk, mean / log-std heads, action =
    (sigmoid(tanh(z)[0]) * max_lin_vel, tanh(tanh(z)[1]) * max_ang_vel) with z ~ N(mean, std) -- the double squashing is
    the reference's (sac.py:97-101).  Attribute names match so its `torch.save(state_dict)` checkpoints load."""

    def __init__(self, num_inputs: int, num_actions: int = 2, hidden_size: int = 256, max_lin_vel: float = MAX_LIN_VEL,
                 max_ang_vel: float = MAX_ANG_VEL, init_w: float = 3e-3, log_std_min: float = -20.0, log_std_max: float = 2.0):
        super().__init__()
        self.log_std_min, self.log_std_max = log_std_min, log_std_max
        self.linear1 = nn.Linear(num_inputs, hidden_size)
        self.linear2 = nn.Linear(hidden_size, hidden_size)
        self.mean_linear = nn.Linear(hidden_size, num_actions)
        self.log_std_linear = nn.Linear(hidden_size, num_actions)
        for lin in (self.mean_linear, self.log_std_linear):
            lin.weight.data.uniform_(-init_w, init_w)
            lin.bias.data.uniform_(-init_w, init_w)
        self.max_lin_vel, self.max_ang_vel = max_lin_vel, max_ang_vel

    def heads(self, state: torch.Tensor):
        x = F.relu(self.linear1(state))
        x = F.relu(self.linear2(x))
        return self.mean_linear(x), torch.clamp(self.log_std_linear(x), self.log_std_min, self.log_std_max)

    def forward(self, state: torch.Tensor, generator: torch.Generator | None = None, deterministic: bool = False) -> torch.Tensor:
        """A sampled action per row (sac.py:92-103 `get_action`); deterministic=True uses z = mean."""
        mean, log_std = self.heads(state)
        z = mean if deterministic else mean + log_std.exp() * torch.randn(mean.shape, device=mean.device, generator=generator)
        a = torch.tanh(z)
        return torch.stack([torch.sigmoid(a[:, 0]) * self.max_lin_vel, torch.tanh(a[:, 1]) * self.max_ang_vel], dim=1)


class TD3Learner:
    """Agent.learn (TD3:225-285): clipped double-Q targets with target-policy smoothing, delayed actor update,
    Polyak averaging.  Hyper-parameters default to the reference's (configs/td3.yaml, TD3DRV:62-72)."""

    def __init__(self, obs_dim: int, device: torch.device, hidden: int = 256, actor_lr: float = 3e-4,
                 critic_lr: float = 3e-4, gamma: float = 0.99, tau: float = 0.005, noise_std: float = 0.2,
                 noise_clip: float = 0.5, policy_update: int = 2, capturable: bool = False):
        self.actor = TD3Actor(obs_dim, 2, hidden).to(device)
        self.critic1, self.critic2 = TD3Critic(obs_dim, 2, hidden).to(device), TD3Critic(obs_dim, 2, hidden).to(device)
        self.t_actor, self.t_critic1, self.t_critic2 = (copy.deepcopy(m) for m in (self.actor, self.critic1, self.critic2))
        # capturable=True keeps Adam's step counters on the device so that a whole update can live in a CUDA graph
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=actor_lr, capturable=capturable)
        self.opt_c1 = torch.optim.Adam(self.critic1.parameters(), lr=critic_lr, capturable=capturable)
        self.opt_c2 = torch.optim.Adam(self.critic2.parameters(), lr=critic_lr, capturable=capturable)
        self.gamma, self.tau, self.noise_std, self.noise_clip, self.policy_update = gamma, tau, noise_std, noise_clip, policy_update
        self.updates = 0

    def _soft(self, net, target):
        with torch.no_grad():
            for p, tp in zip(net.parameters(), target.parameters()):
                tp.mul_(1.0 - self.tau).add_(p, alpha=self.tau)

    def learn(self, batch) -> Dict[str, torch.Tensor]:
        state, action, reward, next_state, done = batch
        with torch.no_grad():
            na = self.t_actor(next_state)
            noise = (torch.randn_like(na) * self.noise_std).clamp(-self.noise_clip, self.noise_clip)
            na = na + noise                       # TD3:247-250: the smoothed target action is NOT clipped to the action box
            tq = torch.min(self.t_critic1(next_state, na), self.t_critic2(next_state, na))
            target = reward + (1.0 - done) * self.gamma * tq
        l1 = F.mse_loss(self.critic1(state, action), target)
        l2 = F.mse_loss(self.critic2(state, action), target)
        self.opt_c1.zero_grad(set_to_none=True); l1.backward(); self.opt_c1.step()
        self.opt_c2.zero_grad(set_to_none=True); l2.backward(); self.opt_c2.step()
        out = {"critic1": l1.detach(), "critic2": l2.detach()}      # device scalars: no host sync in the update
        self.updates += 1
        if self.updates % self.policy_update == 0:
            la = -self.critic1(state, self.actor(state)).mean()
            self.opt_actor.zero_grad(set_to_none=True); la.backward(); self.opt_actor.step()
            self._soft(self.actor, self.t_actor)
            self._soft(self.critic1, self.t_critic1)
            self._soft(self.critic2, self.t_critic2)
            out["actor"] = la.detach()
        return out


@torch.no_grad()
def collect(env, actor: nn.Module, steps: int, sigma: float = 0.0, replay: ReplayRing | None = None,
            generator: torch.Generator | None = None, learner: "TD3Learner | None" = None, learn_every: int = 0,
            batch_size: int = 256) -> Dict[str, float]:
    """Run `steps` vectorised env steps of `env` (a CrowdNavVecEnv with auto-reset) under `actor`.

    Nothing in the loop synchronises with the host: episode statistics are accumulated on the device and read once
    at the end.  With `learner` and `learn_every` > 0 a TD3 update on a replay mini-batch follows every
    `learn_every`-th env step (TD3DRV:128-133: the reference learns once per env step).

    Returns episode statistics in the reference's terms (UTL:53-64) over the episodes that FINISHED inside the
    window -- episodes finished, successes, mean undiscounted return and length -- plus, separately, the episodes the
    window cut off (`censored`, with their partial mean return / length): a fixed window over-represents short
    episodes, so success rates from `collect` are not comparable with evaluate()'s per-world quota."""
    E, dev = env.E, env.device
    obs = env.obs.clone()
    ret = torch.zeros(E, device=dev)
    length = torch.zeros(E, device=dev)
    acc = torch.zeros(4, dtype=torch.float64, device=dev)       # episodes, successes, sum of returns, sum of lengths
    for t in range(steps):
        a = actor(obs)
        a = explore(a, sigma, generator) if sigma > 0 else a.contiguous()
        nobs, r, d = env.step(a)
        if replay is not None:
            replay.add_batch(obs, a, r, nobs, d)
        live = d != 2
        ret += torch.where(live, r, torch.zeros_like(r))
        length += live.float()
        ended = d == 1
        succ = env.counters()[:, 0] == 1                         # one tiny kernel, no read-back
        e64 = ended.to(torch.float64)
        acc += torch.stack([e64.sum(), (ended & succ).to(torch.float64).sum(), (ret.double() * e64).sum(),
                            (length.double() * e64).sum()])
        ret.masked_fill_(ended, 0.0)
        length.masked_fill_(ended, 0.0)
        obs = nobs.clone()
        if learner is not None and learn_every > 0 and replay is not None and (t + 1) % learn_every == 0:
            with torch.enable_grad():
                learner.learn(replay.sample(batch_size, generator))
    running = length > 0
    tail = torch.stack([running.double().sum(), (ret.double() * running).sum(), (length.double() * running).sum()])
    n_eps, n_succ, sum_ret, sum_len = (float(x) for x in acc.cpu())
    n_cens, c_ret, c_len = (float(x) for x in tail.cpu())
    return {"episodes": int(n_eps), "successes": int(n_succ), "success_rate": n_succ / max(n_eps, 1.0),
            "mean_return": sum_ret / max(n_eps, 1.0), "mean_length": sum_len / max(n_eps, 1.0),
            "censored": int(n_cens), "censored_mean_return": c_ret / max(n_cens, 1.0),
            "censored_mean_length": c_len / max(n_cens, 1.0)}


class GraphedCollector:
    """The rollout loop -- policy forward + exploration noise -> env step -> replay append -> episode statistics (->
    TD3 update) -- captured ONCE as a CUDA graph and replayed: possible because nothing in the loop reads anything
    back to the host (ReplayRing keeps its cursor on the device, the env step is a plain kernel launch on the capture
    stream, Adam runs `capturable`).  One replay = `policy_update` env steps, so that the delayed actor update
    (TD3:271) is baked in at the right cadence.  Statistics as in collect()."""

    def __init__(self, env, actor: nn.Module, sigma: float = 1.0, replay: ReplayRing | None = None,
                 learner: "TD3Learner | None" = None, batch_size: int = 256):
        self.env, self.actor, self.sigma, self.replay, self.learner, self.batch = env, actor, sigma, replay, learner, batch_size
        dev = env.device
        self.obs = env.obs.clone()
        self.ret = torch.zeros(env.E, device=dev)
        self.length = torch.zeros(env.E, device=dev)
        self.acc = torch.zeros(4, dtype=torch.float64, device=dev)
        self.steps_per_replay = learner.policy_update if learner is not None else 1
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                           # warm-up outside the capture (allocator, cuBLAS, Adam state)
            for _ in range(3 * self.steps_per_replay):
                self._iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            for _ in range(self.steps_per_replay):
                self._iteration()

    def _iteration(self):
        env = self.env
        with torch.no_grad():
            a = self.actor(self.obs)
            a = explore(a, self.sigma) if self.sigma > 0 else a.contiguous()
            nobs, r, d = env.step(a)
            if self.replay is not None:
                self.replay.add_batch(self.obs, a, r, nobs, d)
            live = d != 2
            self.ret += torch.where(live, r, torch.zeros_like(r))
            self.length += live.float()
            ended = d == 1
            succ = env.counters()[:, 0] == 1
            e64 = ended.to(torch.float64)
            self.acc += torch.stack([e64.sum(), (ended & succ).to(torch.float64).sum(), (self.ret.double() * e64).sum(),
                                     (self.length.double() * e64).sum()])
            self.ret.masked_fill_(ended, 0.0)
            self.length.masked_fill_(ended, 0.0)
            self.obs.copy_(nobs)
        if self.learner is not None:
            self.learner.learn(self.replay.sample(self.batch))

    def run(self, steps: int) -> int:
        """Replay the graph until at least `steps` env steps have been taken; returns the number taken."""
        n = -(-steps // self.steps_per_replay)
        for _ in range(n):
            self.graph.replay()
        if self.learner is not None:
            self.learner.updates += n * self.steps_per_replay     # (learn() ran inside the graph, not in Python)
        return n * self.steps_per_replay

    def stats(self) -> Dict[str, float]:
        n_eps, n_succ, sum_ret, sum_len = (float(x) for x in self.acc.cpu())
        return {"episodes": int(n_eps), "successes": int(n_succ), "success_rate": n_succ / max(n_eps, 1.0),
                "mean_return": sum_ret / max(n_eps, 1.0), "mean_length": sum_len / max(n_eps, 1.0)}


def rollout_throughput(env, policy: nn.Module, steps: int, warmup: int = 5, learn: bool = False, sigma: float = 1.0,
                       capacity: int = 1 << 18, batch_size: int = 256, graph: bool = False) -> Dict[str, float]:
    """env-steps/s of the whole rollout loop -- policy forward (torch) -> env step (CUDA library) -> replay append
    (-> one TD3 update per env step, TD3DRV:128-133) -- timed on the device with one event pair around `steps`
    iterations; no host synchronisation inside the loop.  graph=True replays the loop as a CUDA graph
    (GraphedCollector) instead of issuing it eagerly."""
    dev = env.device
    replay = ReplayRing(capacity, env.D, dev)
    learner = TD3Learner(env.D, dev, capturable=graph) if learn else None
    if learner is not None:
        policy = learner.actor
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        gc = GraphedCollector(env, policy, sigma, replay, learner, batch_size)
        gc.run(warmup)
        torch.cuda.synchronize(dev)
        e0.record()
        steps = gc.run(steps)
        e1.record()
    else:
        kw = dict(sigma=sigma, replay=replay, learner=learner, learn_every=1 if learn else 0, batch_size=batch_size)
        collect(env, policy, warmup, **kw)
        torch.cuda.synchronize(dev)
        e0.record()
        collect(env, policy, steps, **kw)
        e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    return {"value": env.E * steps / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms / steps, "steps": steps,
            "learn": bool(learn), "updates": learner.updates if learner else 0, "replay_rows": len(replay),
            "how": "CUDA graph of the loop, replayed" if graph else "eager PyTorch"}
