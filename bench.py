#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused crowd-navigation step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

One "step" = one pass of the hot path (cn_step: one fused kernel) over one
batch of worlds.  Default workload is BASELINE.json configs[1] (c2): 4096
worlds per GPU, 20 pedestrians, 360 LiDAR samples, K = 8; weak scaling (each
added GPU brings its own 4096 worlds and every rank ends each step holding all
rows: the observation all-gather is fused into the step kernel).  Prints ONE
JSON line (rank 0).

Timing -- ONE protocol for every N: W >= 3 warm-up steps, then exactly K steps
captured in ONE CUDA graph (for N > 1 with the fused gather inside the step
kernels and one final arrival wait) bracketed by ONE event pair on the
launching stream; the steps go round-robin over R independent replicas of the
rank's batch whose total footprint exceeds 2x the 126 MB L2, so every launch
finds its state cold.  Max over ranks.  `per_step_events` keeps the other
protocol (one event pair per step, 256 MiB L2 flush between steps).

Keys beyond the base contract:
  roofline        the step kernel: algorithmic bytes (SURVEY 8d formula) / event time vs measured HBM peak
  roofline_c3     the same for BASELINE configs[2] (16384 worlds), the config the roofline capture is named on (N = 1)
  cpu_baseline    the CPU oracle (a port of the same algorithm) on all host cores, bounded sample
  e2e             same metric through the public host-buffer API (pinned H2D of actions, D2H of obs/reward/done)
  value_l2_warm   K steps back to back from ONE cn_step_n call on one batch (state stays in L2, no graph)
  replicas_only   N > 1: the same graph without any gather (kernel scaling alone)
  gather_verified N > 1: every rank recomputed a FOREIGN shard locally and found its rows in the gathered buffer
  configs3        N = 8: BASELINE configs[3] (65536 worlds, mixed behaviours) through the same protocol
  rollout_td3     N = 1: policy forward + env step + replay append (+ one TD3 update per step), env-steps/s
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"c1": 0, "c2": 1, "c3": 2, "c4": 3, "c5": 4}
WORKLOAD_DESC = {
    "c2": "BASELINE configs[1]: 4096 envs/GPU, 20 pedestrians, 360 LiDAR samples (359 rays), K=8, 5 m room, "
          "U(-0.2,0.2)^2 crowd resampled every 1.5 s, auto-reset",
    "c3": "BASELINE configs[2]: 16384 envs/GPU, 20 pedestrians, 360 samples, K=8 (HBM-roofline capture)",
    "c4": "BASELINE configs[3]: 8192 envs/GPU (65536 on 8), 20 pedestrians, mixed random/towards/crossing, 360 samples, K=8",
    "c5": "BASELINE configs[4]: 16384 envs/GPU, 50 pedestrians, 721 samples (720 rays), K=16",
    "c1": "BASELINE configs[0]: 1 env, 5 pedestrians, 37 samples, K=3",
}
PER_GPU_ENVS = {"c1": 1, "c2": 4096, "c3": 16384, "c4": 8192, "c5": 16384}
L2_BYTES = 126e6


def algorithmic_bytes_per_env_step(n_peds: int, n_samples: int, k: int) -> int:
    """4*[2*16 + 2*8*N + 2 + ((R-1)+7+4K) + 1] + 1."""
    return 4 * (2 * 16 + 2 * 8 * n_peds + 2 + ((n_samples - 1) + 7 + 4 * k) + 1) + 1


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload: str):
    """dram__bytes_read+write per launch of the step kernel from the committed `ncu --set full` capture of this
    workload (profiles/traffic.json names the capture each figure comes from); None when there is none.  It is a
    RECORDED figure of that capture, not a measurement of this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(workload), t.get("_source", {}).get(workload)
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the bench runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index: int):
        self.samples = []
        self._stop = threading.Event()
        self._index = index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self._index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                if float(s[6]) > 0:
                    sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------- the CPU arm
def _oracle_rate(wl: str, risk_faithful: bool, min_seconds: float, min_steps: int, n_envs_cap: int = 4096, threads: int = 0):
    """env-steps/s of the oracle port (oracle/cn_oracle.c, OpenMP over worlds, all host threads) on a bounded sample
    of workload `wl`: the OpenMP pool is warmed first, then at least `min_steps` steps and at least `min_seconds`."""
    import numpy as np
    from crowdnav_b200.config import baseline_config
    from oracle.oracle import OracleEnv
    cores = threads if threads > 0 else (os.cpu_count() or 1)
    n = min(PER_GPU_ENVS[wl], n_envs_cap)
    cfg = baseline_config(WORKLOADS[wl], n_envs=n)
    if risk_faithful:
        cfg.flags |= 8                       # CN_FLAG_RISK_FAITHFUL: the same arm with the reference's own perception block
    env = OracleEnv(cfg, threads=cores)
    env.reset()
    rng = np.random.default_rng(0)
    acts = [np.stack([rng.uniform(0, 0.22, n), rng.uniform(-2, 2, n)], 1).astype(np.float32) for _ in range(8)]
    t0 = time.perf_counter()                 # warm-up: thread pool spin-up, page faults, branch predictors (>= 0.5 s)
    i = 0
    while i < 8 or time.perf_counter() - t0 < 0.5:
        env.step(acts[i % 8])
        i += 1
    steps, t0 = 0, time.perf_counter()
    while steps < min_steps or time.perf_counter() - t0 < min_seconds:
        env.step(acts[steps % 8])
        steps += 1
    dt = time.perf_counter() - t0
    return n * steps / dt, cores, n, steps, dt


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.
    The reference's own loop (ROS + Gazebo, Python 2) cannot run here and is
    real-time-locked at <= 6.67 steps/s (ENV:1201), so the timed arm is the
    oracle port (oracle/cn_oracle.c, OpenMP over worlds, all host threads).
    It is timed like the `cpu_baseline` leg of the GPU arm: warmed pool, at least
    --steps steps and at least 3 s, so that the two agree on the same box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    value, cores, n, steps, dt = _oracle_rate(wl, getattr(args, "risk_faithful", False), 3.0, args.steps)
    line = {
        "impl": "reference", "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "timed_steps": steps, "warmup": max(args.warmup, 8), "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[wl], "sample": "%d worlds per step" % n},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d worlds (%.1f s) of the same workload, OpenMP over worlds, pool warmed"
                                   % (steps, n, dt)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference's own Gazebo+ROS loop is not runnable here; its logged rate is 5.5-8.0 env-steps/s "
                "(BASELINE.md section 2) and it is capped at 6.67/s by time.sleep(0.15)",
    }
    _emit(line)


def cpu_baseline(wl: str, risk_faithful: bool = False):
    """The oracle port on the box's host cores: bounded sample of the same workload (about 10 s)."""
    value, cores, n, steps, dt = _oracle_rate(wl, risk_faithful, 10.0, 20)
    # ... and the same port on ONE core (a single-core figure next to the all-cores one)
    v1, _, n1, s1, dt1 = _oracle_rate(wl, risk_faithful, 2.0, 4, n_envs_cap=1024, threads=1)
    from oracle.oracle import lib as _oracle_lib
    _oracle_lib().orc_set_threads(os.cpu_count() or 1)      # whoever uses the oracle after this gets all cores back
    return {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d steps x %d worlds (%.1f s), oracle/cn_oracle.c with OpenMP over worlds, pool warmed" % (steps, n, dt),
            "one_core": {"value": v1, "unit": "env-steps/s", "cores": 1,
                         "sample": "%d steps x %d worlds (%.1f s), the same port on one thread" % (s1, n1, dt1)}}


_REAL_STDOUT = None


def _claim_stdout():
    """Libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line on stdout.  Point fd 1 at
    stderr for the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# ----------------------------------------------------------------------------------------- the GPU arm
class Bench:
    """Everything one rank needs to time one workload."""

    def __init__(self, args, wl, per_gpu, gather_mode):
        import torch
        import torch.distributed as dist
        from crowdnav_b200.config import CN_FLAG_RISK_FAITHFUL, baseline_config
        from crowdnav_b200.sharded import ShardedVecEnv
        from crowdnav_b200.vec_env import CrowdNavVecEnv
        self.torch, self.dist, self.args, self.wl = torch, dist, args, wl
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local_rank)
        self.cfg_global = baseline_config(WORKLOADS[wl], n_envs=per_gpu * self.world, auto_reset=True)
        if args.risk_faithful:
            self.cfg_global.flags |= CN_FLAG_RISK_FAITHFUL
        self.gather_mode = gather_mode if self.world > 1 else "none"
        mk = lambda c, o: CrowdNavVecEnv(c, device=self.local_rank, obs_out=o)
        self.fallback = None
        try:
            self.senv = ShardedVecEnv(self.cfg_global, mk, self.dev, gather=self.gather_mode if self.world > 1 else "collective")
        except Exception as exc:                       # symmetric memory / multicast unavailable: say so, fall back
            if not self.gather_mode.startswith("fused"):
                raise
            nxt = "fused" if self.gather_mode in ("fused_mc", "fused_async", "fused_async16") else "collective"
            self.fallback = "%s unavailable (%s); using %s" % (self.gather_mode, str(exc).splitlines()[0][:160], nxt)
            if self.rank == 0:
                print(self.fallback, file=sys.stderr)
            self.gather_mode = nxt
            self.senv = ShardedVecEnv(self.cfg_global, mk, self.dev, gather=nxt)
        self.env = self.senv.env
        self.E_local, self.E_total, self.D = self.env.E, self.cfg_global.n_envs, self.env.D
        self.bytes_per_env = algorithmic_bytes_per_env_step(self.cfg_global.n_peds, self.cfg_global.n_samples,
                                                            self.cfg_global.k_obstacles)
        # actions: the reference's exploration policy, a TD3 actor (random init: no checkpoint travels to the box) +
        # N(0,1) noise, clipped (TD3:81-106, 196-223); configs[2] ("SAC rollout") samples the reference's SAC policy
        from crowdnav_b200.rollout import SACActor, TD3Actor, explore
        torch.manual_seed(1234 + self.rank)
        self.policy_kind = "SAC" if wl == "c3" else "TD3"
        self.actor = (SACActor(self.D) if wl == "c3" else TD3Actor(self.D)).to(self.dev)
        self._explore = explore
        self.senv.reset()
        self.ring = []
        for i in range(16):                      # closed-loop warm start: 16 policy steps fill the action ring
            a = self.policy(self.senv.obs_local)
            self.ring.append(a)
            self.senv.step_local(a)
            self.senv.gather()
        self.senv.wait_gathered()
        torch.cuda.synchronize()
        self.reps = None

    def policy(self, obs):
        torch = self.torch
        with torch.no_grad():
            if self.policy_kind == "SAC":
                return self.actor(obs).contiguous()          # a sample of the stochastic policy (sac.py:92-103)
            return self._explore(self.actor(obs), 1.0)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def replicas(self):
        """R independent copies of this rank's batch, > 2x L2 in total, each advanced 16 steps."""
        if self.reps is None:
            from crowdnav_b200.vec_env import CrowdNavVecEnv
            per = self.E_local * (64 + 32 * self.cfg_global.n_peds + 4 * self.D + 5)
            R = max(2, int(math.ceil(2.4 * L2_BYTES / per)))
            reps = [self.env]
            for _ in range(R - 1):
                r_ = CrowdNavVecEnv(self.senv.cfg_local, device=self.local_rank)
                r_.reset()
                for i in range(16):
                    r_.step(self.ring[i])
                reps.append(r_)
            self.torch.cuda.synchronize()
            self.reps, self.rep_mb = reps, per / 1e6
        return self.reps

    def graph_timed(self, K, warmup, gather=True):
        """K steps as ONE CUDA graph inside ONE event pair, cold state by rotation over the replicas.  With
        gather=False (N > 1) the steps are plain cn_step launches into the local buffers: replicas only."""
        torch = self.torch
        reps = self.replicas()
        R = len(reps)
        senv = self.senv
        scratch = [torch.zeros((self.E_local, self.D), dtype=torch.float32, device=self.dev) for _ in range(3)] \
            if (not gather and self.world > 1) else None

        def one(i):
            if self.world == 1:
                reps[i % R].step(self.ring[i % 16])          # every replica keeps its own observation buffer
            elif gather:
                senv.step_local(self.ring[i % 16], env=reps[i % R])
                senv.gather()
            else:
                reps[i % R].obs = scratch[i % 3]
                reps[i % R].step(self.ring[i % 16])

        for i in range(warmup):
            one(i)
        if gather:
            senv.wait_gathered()
        self.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(K):
                one(i)
            if gather:
                senv.wait_gathered()             # every rank holds every row of the last step when the graph ends
        self.barrier()
        graph.replay()                           # one untimed pass: K more warm-up steps
        self.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        graph.replay()                           # EXACTLY K step launches (+ one arrival-wait kernel for N > 1)
        g1.record()
        self.barrier()
        ms = g0.elapsed_time(g1)
        del graph
        return ms, {"replicas": R, "replica_mb": self.rep_mb}

    def events_timed(self, K, warmup, do_flush=True):
        """One event pair per step, 256 MiB L2 flush between steps (outside the pairs)."""
        torch = self.torch
        senv = self.senv
        if not hasattr(self, "_flush"):
            self._flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)
        tot = kern = 0.0
        for phase, n in (("warm", warmup), ("timed", K)):
            evs = []
            for i in range(n):
                if do_flush:
                    self._flush.fill_(float(i))
                s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                s0.record()
                senv.step_local(self.ring[i % 16])
                s1.record()
                senv.gather()
                s2.record()
                evs.append((s0, s1, s2))
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            senv.wait_gathered()
            t1.record()
            self.barrier()
            if phase == "timed":
                tot = sum(a.elapsed_time(c) for a, b, c in evs) + t0.elapsed_time(t1)
                kern = sum(a.elapsed_time(b) for a, b, c in evs)
        return tot, kern

    def step_n_timed(self, K, warmup):
        """K steps enqueued by ONE cn_step_n call (C loop of launches, no graph, no Python per step), one batch."""
        torch = self.torch
        acts = torch.stack(self.ring[:16]).repeat((K + 15) // 16, 1, 1)[:K].contiguous()
        rew = torch.empty((K, self.E_local), dtype=torch.float32, device=self.dev)
        done = torch.empty((K, self.E_local), dtype=torch.uint8, device=self.dev)
        self.env.obs = self.senv.obs_local
        self.env.step_n(acts[:max(warmup, 1)].contiguous(), rew[:max(warmup, 1)], done[:max(warmup, 1)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.env.step_n(acts, rew, done)
        e1.record()
        torch.cuda.synchronize()
        stepn_ms = e0.elapsed_time(e1)
        # the same K steps as a CUDA graph owned by the LIBRARY (cn_graph_create / cn_graph_launch: no torch capture)
        g = self.env.make_graph(acts)
        g.launch()
        torch.cuda.synchronize()
        e0.record()
        g.launch()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1)
        g.close()
        return stepn_ms, graph_ms

    def max_over_ranks(self, vals):
        torch = self.torch
        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    # -- the all-gather's own correctness: single-GPU batch == gathered rows
    def verify_gather(self, steps=6):
        """Every rank steps, with fresh handles and the SAME deterministic actions, (a) its own shard through the
        gather under test and (b) a FOREIGN shard (the next rank's env ids) locally, and compares -- bit for bit, every
        step -- the foreign rows that arrived in its gathered buffer with what it computed itself."""
        torch, dist = self.torch, self.dist
        from crowdnav_b200.sharded import local_config
        from crowdnav_b200.vec_env import CrowdNavVecEnv
        senv = self.senv
        frank = (self.rank + 1) % self.world
        mine = CrowdNavVecEnv(senv.cfg_local, device=self.local_rank)
        other = CrowdNavVecEnv(local_config(self.cfg_global, frank, self.world), device=self.local_rank)
        flo, fhi = frank * self.E_local, (frank + 1) * self.E_local
        senv.wait_gathered()
        self.barrier()
        mine.reset()
        other.reset()
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(4242)                                   # the same stream of [E_total, 2] batches on every rank
        ok = True
        for s in range(steps):
            u = torch.rand((self.E_total, 2), device=self.dev, generator=gen)
            act = torch.stack([u[:, 0] * 0.22, u[:, 1] * 4.0 - 2.0], 1).contiguous()
            senv.step_local(act[senv.lo:senv.hi].contiguous(), env=mine)
            senv.gather()
            senv.wait_gathered()
            fobs, _, _ = other.step(act[flo:fhi].contiguous())
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(senv.obs_all[flo:fhi].view(torch.int32), fobs.view(torch.int32))) \
                and bool(torch.equal(senv.obs_all[senv.lo:senv.hi], mine.obs))
            self.barrier()                                      # nobody runs ahead while a peer still compares
        if senv.async_mode:
            # ... and the pipelined path proper: rows forwarded by the NEXT step's kernel (and, in the 16-bit format, rebuilt
            # by the one after), no flush in between
            hist = []
            for s in range(steps + senv.lag):
                u = torch.rand((self.E_total, 2), device=self.dev, generator=gen)
                act = torch.stack([u[:, 0] * 0.22, u[:, 1] * 4.0 - 2.0], 1).contiguous()
                senv.step_local(act[senv.lo:senv.hi].contiguous(), env=mine)
                got = senv.wait_pushed()
                fobs, _, _ = other.step(act[flo:fhi].contiguous())
                torch.cuda.synchronize()
                hist.append(fobs.clone())
                k = len(hist) - 1 - senv.lag
                if k >= 0:
                    ok = ok and got is not None and bool(torch.equal(got[flo:fhi].view(torch.int32), hist[k].view(torch.int32)))
                self.barrier()
            ok = ok and bool(torch.equal(senv.wait_gathered()[flo:fhi].view(torch.int32), hist[-1].view(torch.int32)))
            self.barrier()
        timeouts = (mine.gather_timeouts + senv.env.gather_timeouts) if senv.fused else 0
        flag = torch.tensor([1 if (ok and timeouts == 0) else 0], dtype=torch.int64, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        mine.close()
        other.close()
        return bool(flag.item()), {"steps": steps, "rows_compared_per_rank": 2 * self.E_local * steps,
                                   "foreign_shard": "rank + 1", "compared": "bit patterns",
                                   "gather_timeouts": int(timeouts)}


def measure(args, wl, per_gpu, gather_mode, K, warmup, full=True):
    """Time workload `wl`; returns the dict of results (rank 0 fills the JSON line from it)."""
    B = Bench(args, wl, per_gpu, gather_mode)
    world = B.world
    out = {"bench": B}
    graph_ms, ginfo = B.graph_timed(K, warmup, gather=True)
    out["graph_ms"], out["ginfo"] = B.max_over_ranks([graph_ms])[0], ginfo
    # step kernels + (fused: the final wait) / (async: flush + wait) / (async16: flush + wait + decode of the last two steps)
    out["launches_per_rank"] = K + {"fused": 1, "fused_mc": 1, "fused_async": 2, "fused_async16": 4}.get(
        B.gather_mode if world > 1 else "", 0)
    if world > 1:
        none_ms, _ = B.graph_timed(K, warmup, gather=False)
        out["none_ms"] = B.max_over_ranks([none_ms])[0]
    if full:
        ev_tot, ev_kern = B.events_timed(K, warmup, do_flush=True)
        out["ev_tot"], out["ev_kern"] = B.max_over_ranks([ev_tot, ev_kern])
        if world == 1:
            out["stepn_ms"], out["libgraph_ms"] = B.step_n_timed(K, warmup)
    return out


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "fused_async16", "fused_async", "fused", "fused_mc", "nccl"],
                    help="N>1: 'auto' = 'fused_async' at 2 GPUs (6.5 MB of rows per rank fit under the step's compute as they are), "
                         "'fused_async16' from 3 GPUs on (the step is NVLink-bound: halve the bytes); 'fused_async16' = 'fused_async' with the rows travelling as int16 thousandths (half the NVLink "
                         "bytes, rebuilt bit for bit by a decode kernel behind the arrival wait); 'fused_async' = pipelined fused gather: the kernel of step t+1 forwards the rows of step t to "
                         "every peer (bulk TMA through a staging tile) under its own compute and signals the peers' arrival "
                         "counters; one push-only launch flushes the last step; 'fused' = the step kernel stores its OWN rows "
                         "into every peer at its end and signals -- no other launch per step; 'fused_mc' = the same with "
                         "NVSwitch multicast (multimem.st / multimem.red); 'nccl' = separate in-place ncclAllGather after the kernel")
    ap.add_argument("--risk-faithful", action="store_true",
                    help="CN_FLAG_RISK_FAITHFUL: K block and counters from the reference's own segmentation / tracker "
                         "(cn_faithful_kernel runs behind the step kernel: two launches per step); single GPU or --gather nccl")
    ap.add_argument("--skip-verify", action="store_true", help="diagnostics: do not run the gather self-check")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline_c3 / configs3 / rollout_td3 / verification extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    K = args.steps
    wl = args.workload
    per_gpu = args.envs_per_gpu or PER_GPU_ENVS[wl]
    gather_mode = "none"
    if world > 1:
        gather_mode = {"nccl": "collective", "auto": "fused_async" if world <= 2 else "fused_async16"}.get(args.gather, args.gather)
        if args.risk_faithful:
            gather_mode = "collective"

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    t_wall0 = time.perf_counter()

    M = measure(args, wl, per_gpu, gather_mode, K, warmup, full=True)
    B = M["bench"]
    env, senv = B.env, B.senv
    E_local, E_total, D = B.E_local, B.E_total, B.D
    cfg_global = B.cfg_global

    # --- N > 1: prove the gather (single-GPU batch == gathered rows) before anything else is reported
    gather_verified, gather_check = None, None
    if world > 1 and not args.risk_faithful and not args.skip_verify:
        gather_verified, gather_check = B.verify_gather()

    # --- e2e through the public host-buffer API: per step pinned H2D of the actions, the kernel, D2H of obs / reward /
    # done and a stream synchronise.  N = 1: three variants (see CrowdNavVecEnv.step_host / step_host_pipelined).
    h_actions = [r_.cpu().numpy() for r_ in B.ring]
    e2e_modes = {}
    env.obs = senv.obs_local

    def time_host(fn, tail=None):
        for i in range(warmup):
            fn(h_actions[i % 16])
        if tail:
            tail()
        B.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            fn(h_actions[i % 16])
        if tail:
            tail()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        B.barrier()
        return dt

    if world == 1:
        e2e_modes["copy"] = time_host(lambda a: env.step_host(a, mode="copy"))
        if not args.risk_faithful:
            try:
                e2e_modes["mapped"] = time_host(lambda a: env.step_host(a, mode="mapped"))
            except Exception as exc:                               # e.g. pinned memory not device-addressable
                print("step_host(mode='mapped') unavailable: %s" % exc, file=sys.stderr)
        e2e_modes["pipelined"] = time_host(env.step_host_pipelined, tail=env.flush_host_pipeline)
    else:
        e2e_modes["copy"] = time_host(senv.step_host)
    e2e_modes = {k: B.max_over_ranks([v])[0] for k, v in e2e_modes.items()}

    # --- extras
    extras = {}
    if not args.no_extras and not args.risk_faithful:
        if world == 1 and wl == "c2":
            M3 = measure(args, "c3", PER_GPU_ENVS["c3"], "none", K, warmup, full=False)
            extras["c3"] = M3
        if world == 8 and wl == "c2":
            M4 = measure(args, "c4", PER_GPU_ENVS["c4"], gather_mode, K, warmup, full=False)
            extras["c4"] = M4
            extras["c4_verified"] = M4["bench"].verify_gather(steps=3)
        if world == 1:
            from crowdnav_b200.rollout import rollout_throughput
            n_r = max(20, min(K, 100))
            env.obs = senv.obs_local
            extras["rollout"] = {"policy_only": rollout_throughput(env, B.actor, n_r, learn=False),
                                 "with_td3_update": rollout_throughput(env, B.actor, n_r, learn=True),
                                 "what": "torch policy forward + exploration noise -> cn_step -> device replay append "
                                         "(-> one TD3 update on a 256-row mini-batch per env step, TD3DRV:128-133); "
                                         "PyTorch around the library call, no host synchronisation in the loop; the "
                                         "*_graph entries replay the same loop as one CUDA graph (GraphedCollector)"}
            for key, learn in (("policy_only_graph", False), ("with_td3_update_graph", True)):
                try:
                    extras["rollout"][key] = rollout_throughput(env, B.actor, n_r, learn=learn, graph=True)
                except Exception as exc:                       # a capture problem must not cost the bench line
                    extras["rollout"][key] = {"error": str(exc).splitlines()[0][:200]}
                    torch.cuda.synchronize()

    # keep the GPU busy long enough for the clock sampler to see it under load; every rank must run the SAME
    # number of extra steps (each step holds a cross-rank signal / collective), so rank 0 decides
    remaining = max(0.0, 2.5 - (time.perf_counter() - t_wall0))
    rounds = torch.tensor([int(remaining / max(50 * (M["graph_ms"] / K) * 1e-3, 1e-4)) + 1], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(rounds, src=0)
    for _ in range(min(int(rounds.item()), 2000)):
        for i in range(50):
            senv.step_local(B.ring[i % 16])
            senv.gather()
    senv.wait_gathered()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ms_per_step = M["graph_ms"] / K
        kern_s = ms_per_step * 1e-3              # average launch duration over the timed region (incl. launch gaps)
        ginfo = M["ginfo"]
        l2_text = ("cold by rotation: the K steps go round-robin over %d independent replicas of the rank's batch "
                   "(%.1f MB each, %.0f MB > 2x the 126 MB L2); the K launches are one CUDA graph inside one event pair"
                   % (ginfo["replicas"], ginfo["replica_mb"], ginfo["replicas"] * ginfo["replica_mb"]))
        achieved = E_local * B.bytes_per_env / kern_s / 1e9
        traffic, traffic_src = recorded_traffic(wl)
        gather_text = {
            "none": "single GPU",
            "fused": "env-id sharding x%d; obs all-gather fused into the step kernel: every CTA stores its tile of rows into "
                     "every peer's symmetric-memory buffer (16-byte stores over NVLink, peer order rotated per rank and per "
                     "CTA) and signals the peers' arrival counters (red.release.sys); 3 rotating buffers, device-side "
                     "step counting, no other launch per step" % world,
            "fused_mc": "env-id sharding x%d; obs all-gather fused into the step kernel with NVSwitch multicast: one "
                        "multimem.st per 16 bytes reaches every rank's buffer, signal by multimem.red; 3 rotating buffers" % world,
            "fused_async": "env-id sharding x%d; obs all-gather fused into the step kernels and PIPELINED: the kernel of step t+1 "
                           "forwards the rows of step t (bulk TMA load into a staging tile, bulk TMA stores into every peer's "
                           "symmetric-memory buffer, peer order rotated per rank and per CTA) under its own compute and signals "
                           "the peers' arrival counters; the timed graph ends with one push-only launch for the last step and "
                           "the arrival wait, so all K steps' rows have landed on every rank inside the timed region" % world,
            "fused_async16": "env-id sharding x%d; obs all-gather fused into the step kernels, PIPELINED, 16-bit wire format: the step "
                             "kernel also writes its rows as int16 thousandths (every row value is a whole number of thousandths), "
                             "the kernel of step t+1 forwards step t's int16 rows (bulk TMA through a staging tile into every peer's "
                             "symmetric-memory wire buffer, peer order rotated per rank and per CTA) under its own compute and "
                             "signals the peers' arrival counters, and the kernel of step t+2 rebuilds the peers' fp32 rows of "
                             "step t bit for bit while its state tile loads -- one launch per step; the timed graph ends with the "
                             "push-only launch, the arrival wait and the decode of the last two steps, so every rank holds all K "
                             "steps' fp32 rows inside the timed region" % world,
            "collective": "env-id sharding x%d; one in-place ncclAllGather per step" % world}[B.gather_mode]
        line = {
            "metric": "env-steps/s", "value": E_total / (ms_per_step * 1e-3), "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[wl], "envs_per_gpu": E_local, "envs_total": E_total,
                       "n_peds": cfg_global.n_peds, "n_samples": cfg_global.n_samples, "k_obstacles": cfg_global.k_obstacles,
                       "obs_dim": D, "parallelism": gather_text, "l2": l2_text,
                       "timing": "one CUDA graph of K steps inside one event pair, max over ranks (the same protocol for every N)",
                       "risk_block": ("faithful: the reference's own segmentation / tracker in float64, cn_faithful_kernel "
                                      "behind the step kernel (2 launches per step)" if args.risk_faithful else
                                      "intended: ideal association inside the step kernel"),
                       "actions": "ring of 16 batches from a random-init %s actor%s" % (
                           B.policy_kind, " + N(0,1) exploration noise, clipped" if B.policy_kind == "TD3" else
                           " (samples of the stochastic policy)")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_recorded_from": traffic_src, "peak_source": peak_src,
                         "kernel": "%s<step>, %d worlds per CTA" % (env.kernel_name, env.kernel_tile), "kernel_us": kern_s * 1e6,
                         "algorithmic_bytes_per_env_step": B.bytes_per_env, "envs_per_launch": E_local},
            "per_step_events": {"kernel_us": M["ev_kern"] / K * 1e3, "ms_per_step": M["ev_tot"] / K,
                                "value": E_total / (M["ev_tot"] / K * 1e-3),
                                "frac": E_local * B.bytes_per_env / (M["ev_kern"] / K * 1e-3) / 1e9 / peak,
                                "l2": "flushed between timed steps (256 MiB write, outside the event pairs); one event pair "
                                      "per step, which itself costs 2.6 us (profiles/tools/launch_overhead.py)"},
            "gpu_launches": int(M["launches_per_rank"] * world),
            "clocks": clocks,
        }
        if "stepn_ms" in M:
            sn = M["stepn_ms"] / K
            line["value_l2_warm"] = E_total / (sn * 1e-3)
            lg = M["libgraph_ms"] / K
            line["value_l2_warm_lib_graph"] = E_total / (lg * 1e-3)
            line["roofline"]["l2_warm_lib_graph"] = {"kernel_us": lg * 1e3, "frac": E_local * B.bytes_per_env / (lg * 1e-3) / 1e9 / peak,
                                                     "how": "the same K launches as ONE graph built by cn_graph_create and replayed by "
                                                            "cn_graph_launch (C ABI, no torch capture), one batch, no flush"}
            line["roofline"]["l2_warm"] = {"kernel_us": sn * 1e3, "achieved": E_local * B.bytes_per_env / (sn * 1e-3) / 1e9,
                                           "frac": E_local * B.bytes_per_env / (sn * 1e-3) / 1e9 / peak,
                                           "how": "K launches enqueued by one cn_step_n call (C loop, no graph), one batch, no flush"}
        if "none_ms" in M:
            line["replicas_only"] = {"value": E_total / (M["none_ms"] / K * 1e-3), "ms_per_step": M["none_ms"] / K,
                                     "what": "the same graph-timed steps without any gather (plain cn_step per rank)"}
        if gather_verified is not None:
            line["gather_verified"] = gather_verified
            line["gather_check"] = gather_check
        if B.fallback:
            line["gather_fallback"] = B.fallback
        if world > 1:
            line["multicast_available"] = bool(getattr(senv, "multicast_available", False))
        best = min((k for k in e2e_modes if k != "pipelined"), key=lambda k: e2e_modes[k])
        line["e2e"] = {"value": E_total / (e2e_modes[best] / K), "unit": "env-steps/s",
                       "h2d_bytes_per_step": env.h2d_bytes_per_step * world, "d2h_bytes_per_step": env.d2h_bytes_per_step * world,
                       "ms_per_step": 1e3 * e2e_modes[best] / K,
                       "api": ("CrowdNavVecEnv.step_host(mode=%r) (pinned host buffers; strict: returns the step just issued)" % best)
                              if world == 1 else "ShardedVecEnv.step_host (pinned host buffers, gathered rows certified every step)",
                       "modes": {k: {"value": E_total / (v / K), "ms_per_step": 1e3 * v / K} for k, v in e2e_modes.items()},
                       "modes_doc": {"copy": "H2D actions, kernel, 3 D2H copies by the copy engine, stream sync",
                                     "mapped": "H2D actions, the kernel writes obs / reward / done straight into the pinned host "
                                               "buffers (bulk stores over PCIe), stream sync -- same bytes, no separate copies",
                                     "pipelined": "opt-in, one-step-stale results: D2H of step t on a copy stream under kernel "
                                                  "t+1; NOT the headline (value is the best strict mode)"}}
        if "c3" in extras:
            M3 = extras["c3"]
            B3 = M3["bench"]
            ks = M3["graph_ms"] / K * 1e-3
            t3, t3src = recorded_traffic("c3")
            line["roofline_c3"] = {"bound": "hbm", "achieved": B3.E_local * B3.bytes_per_env / ks / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": B3.E_local * B3.bytes_per_env / ks / 1e9 / peak, "traffic": t3,
                                   "traffic_recorded_from": t3src, "kernel_us": ks * 1e6,
                                   "kernel": "%s<step>, %d worlds per CTA" % (B3.env.kernel_name, B3.env.kernel_tile),
                                   "value": B3.E_total / ks, "workload": WORKLOAD_DESC["c3"], "actions": "random-init SAC policy samples",
                                   "algorithmic_bytes_per_env_step": B3.bytes_per_env, "envs_per_launch": B3.E_local,
                                   "replicas": M3["ginfo"]["replicas"]}
        if "c4" in extras:
            M4 = extras["c4"]
            B4 = M4["bench"]
            line["configs3"] = {"workload": WORKLOAD_DESC["c4"], "value": B4.E_total / (M4["graph_ms"] / K * 1e-3),
                                "ms_per_step": M4["graph_ms"] / K, "envs_total": B4.E_total, "gather": B4.gather_mode,
                                "replicas_only": {"value": B4.E_total / (M4["none_ms"] / K * 1e-3), "ms_per_step": M4["none_ms"] / K},
                                "gather_verified": extras["c4_verified"][0], "gather_check": extras["c4_verified"][1]}
        if "rollout" in extras:
            line["rollout_td3"] = extras["rollout"]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl, args.risk_faithful)
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
