#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused crowd-navigation step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

One "step" = one pass of the hot path (cn_step: one fused kernel) over one
batch of worlds.  Default workload is BASELINE.json configs[1] (c2): 4096
worlds per GPU, 20 pedestrians, 360 LiDAR samples, K = 8; weak scaling (each
added GPU brings its own 4096 worlds, one in-place all-gather of the
observation tensor per step).  Prints ONE JSON line (rank 0).

Timing: W >= 3 warm-up steps, then exactly K steps, each bracketed by CUDA
events on the launching stream; L2 is flushed (256 MiB write) between timed
steps because the c2 working set (12 MB) fits the 126 MB L2; the flush is
outside the event pairs.  Max over ranks.

Keys beyond the base contract:
  roofline      the step kernel: algorithmic bytes (SURVEY 8d formula) / event time vs measured HBM peak
  cpu_baseline  the CPU oracle (a port of the same algorithm) on all host cores, bounded sample
  e2e           same metric through the public host-buffer API (pinned H2D of actions, D2H of obs/reward/done)
  value_l2_warm same steps back to back without the flush (state stays in L2, as in a real rollout)
"""
from __future__ import annotations

import argparse
import json
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
    """dram__bytes_read+write per launch from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


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


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.
    The reference's own loop (ROS + Gazebo, Python 2) cannot run here and is
    real-time-locked at <= 6.67 steps/s (ENV:1201), so the timed arm is the
    oracle port (oracle/cn_oracle.c, OpenMP over worlds, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from crowdnav_b200.config import baseline_config
    from oracle.oracle import OracleEnv
    wl = args.workload
    cores = os.cpu_count() or 1
    n_envs = PER_GPU_ENVS[wl] * args.gpus
    sample_envs = min(n_envs, 4096)          # bounded sample: one step of <= 4096 worlds
    cfg = baseline_config(WORKLOADS[wl], n_envs=sample_envs)
    if getattr(args, "risk_faithful", False):
        cfg.flags |= 8                       # CN_FLAG_RISK_FAITHFUL: the same arm with the reference's own perception block
    env = OracleEnv(cfg, threads=cores)
    env.reset()
    rng = np.random.default_rng(0)
    acts = [np.stack([rng.uniform(0, 0.22, sample_envs), rng.uniform(-2, 2, sample_envs)], 1).astype(np.float32)
            for _ in range(8)]
    for i in range(max(args.warmup, 3)):
        env.step(acts[i % 8])
    t0 = time.perf_counter()
    for i in range(args.steps):
        env.step(acts[i % 8])
    dt = time.perf_counter() - t0
    value = sample_envs * args.steps / dt
    line = {
        "impl": "reference", "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[wl], "sample": "%d worlds per step" % sample_envs},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d worlds of the same workload, OpenMP over worlds" % (args.steps, sample_envs)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference's own Gazebo+ROS loop is not runnable here; its logged rate is 5.5-8.0 env-steps/s "
                "(BASELINE.md section 2) and it is capped at 6.67/s by time.sleep(0.15)",
    }
    _emit(line)


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
    ap.add_argument("--timing", default="graph", choices=["graph", "events"],
                    help="single GPU: 'graph' = the K timed steps are captured in ONE CUDA graph (round-robin over "
                         "replicas of the batch whose total size exceeds 2x L2) and bracketed by one event pair; "
                         "'events' = one event pair per step with a 256 MiB L2 flush in between (always used for N > 1)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N>1: 'fused' = the step kernel stores its rows into every peer's gather buffer (symmetric "
                         "memory, NVLink) + a barrier; 'nccl' = separate in-place ncclAllGather after the kernel")
    ap.add_argument("--risk-faithful", action="store_true",
                    help="CN_FLAG_RISK_FAITHFUL: K block and counters from the reference's own segmentation / tracker "
                         "(cn_faithful_kernel runs behind the step kernel: two launches per step); single GPU or --gather nccl")
    ap.add_argument("--with-policy", action="store_true",
                    help="run the TD3 actor forward (torch) inside each step instead of replaying action batches")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from crowdnav_b200.config import baseline_config
    from crowdnav_b200.sharded import ShardedVecEnv
    from crowdnav_b200.vec_env import CrowdNavVecEnv

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
    cfg_global = baseline_config(WORKLOADS[wl], n_envs=per_gpu * world, auto_reset=True)
    if args.risk_faithful:
        from crowdnav_b200.config import CN_FLAG_RISK_FAITHFUL
        cfg_global.flags |= CN_FLAG_RISK_FAITHFUL
    gather_mode = "none"
    if world > 1:
        gather_mode = "fused" if (args.gather == "fused" and not args.risk_faithful) else "collective"
    try:
        senv = ShardedVecEnv(cfg_global, lambda c, o: CrowdNavVecEnv(c, device=local_rank, obs_out=o), dev,
                             gather=gather_mode if world > 1 else "collective")
    except Exception as exc:                       # symmetric memory unavailable: say so, use the collective
        if gather_mode != "fused":
            raise
        if rank == 0:
            print("fused gather unavailable (%s); using ncclAllGather" % exc, file=sys.stderr)
        gather_mode = "collective"
        senv = ShardedVecEnv(cfg_global, lambda c, o: CrowdNavVecEnv(c, device=local_rank, obs_out=o), dev,
                             gather="collective")
    env = senv.env
    E_local, E_total, D = env.E, cfg_global.n_envs, env.D
    bytes_per_env = algorithmic_bytes_per_env_step(cfg_global.n_peds, cfg_global.n_samples, cfg_global.k_obstacles)

    # --- actions: the reference's exploration policy, a TD3 actor (random init: no
    # checkpoint travels to the box) + N(0,1) noise, clipped (TD3:81-106, 196-223)
    torch.manual_seed(1234 + rank)
    actor = torch.nn.Sequential(torch.nn.Linear(D, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                                torch.nn.Linear(256, 2)).to(dev)

    def policy(obs):
        with torch.no_grad():
            a = actor(obs)
            v = torch.sigmoid(a[:, 0]) * 0.22 + torch.randn(obs.shape[0], device=dev)
            w = torch.tanh(a[:, 1]) * 2.0 + torch.randn(obs.shape[0], device=dev)
            return torch.stack([v.clamp(0.0, 0.22), w.clamp(-2.0, 2.0)], 1).contiguous()

    senv.reset()
    ring = []
    for i in range(16):                      # closed-loop warm start: 16 policy steps fill the action ring
        a = policy(senv.obs_local)
        ring.append(a)
        senv.step(a)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    t_wall0 = time.perf_counter()

    def run_steps(n, timed, do_flush, kernel_only=False):
        """n steps; returns (sum of per-step device ms, sum of kernel-only ms)."""
        evs = []
        for i in range(n):
            if do_flush:
                flush.fill_(float(i))
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            a = policy(senv.obs_local) if args.with_policy else ring[i % 16]
            if world > 1:
                senv.begin_step()
            env.step(a)                 # the kernel (in fused mode it also stores the rows into the peers)
            s1.record()
            s2 = s1                     # single GPU: the step IS the kernel launch, nothing follows it
            if world > 1:
                senv.gather()           # ncclAllGather, or just the cross-rank barrier in fused mode
                s2 = torch.cuda.Event(enable_timing=True)
                s2.record()
            if timed:
                evs.append((s0, s1, s2))
        # pipelined gather (fused mode): the last barriers finish after the last step's events
        tail0, tail1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tail0.record()
        if world > 1:
            senv.wait_gathered()
        tail1.record()
        torch.cuda.synchronize()
        step_ms = sum(a.elapsed_time(c) for a, b, c in evs) + (tail0.elapsed_time(tail1) if timed else 0.0)
        kern_ms = sum(a.elapsed_time(b) for a, b, c in evs)
        return step_ms, kern_ms

    # --- device-resident timing (value): K steps, L2 flushed between steps
    run_steps(warmup, False, True)
    barrier()
    launches0 = env.launch_count
    step_ms, kern_ms = run_steps(K, True, True)
    launches = env.launch_count - launches0
    barrier()
    # --- same, back to back without the flush (state L2-resident)
    run_steps(warmup, False, False)
    barrier()
    warm_ms, warm_kern_ms = run_steps(K, True, False)
    barrier()

    # --- single GPU: the K timed steps as ONE CUDA graph (what a launch-bound rollout loop does), cold in L2 because the
    # steps go round-robin over R replicas of the batch whose total footprint is > 2x L2.  One event pair brackets exactly
    # K launches: no per-step event overhead (2.6 us for two back-to-back records, profiles/tools/launch_overhead.py), no
    # flush kernel inside or next to the timed region.
    graph_ms = None
    graph_info = None
    if world == 1 and args.timing == "graph" and not args.with_policy:
        import math
        per = E_local * (64 + 32 * cfg_global.n_peds + 4 * D + 5)
        R = max(2, int(math.ceil(2.4 * 126e6 / per)))
        reps = [env]
        for _ in range(R - 1):
            r_ = CrowdNavVecEnv(senv.cfg_local, device=local_rank)
            r_.reset()
            for i in range(16):
                r_.step(ring[i])
            reps.append(r_)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(K):
                reps[i % R].step(ring[i % 16])
        for i in range(warmup):
            reps[i % R].step(ring[i % 16])
        graph.replay()                       # one untimed pass: K more warm-up steps
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        graph.replay()                       # EXACTLY K step launches
        g1.record()
        barrier()
        graph_ms = g0.elapsed_time(g1)
        graph_info = {"replicas": R, "replica_mb": per / 1e6}

    # --- e2e through the public host-buffer API
    h_actions = [r_.cpu().numpy() for r_ in ring]
    for i in range(warmup):
        env.step_host(h_actions[i % 16])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        env.step_host(h_actions[i % 16])        # pinned H2D + kernel + D2H + stream sync, every step
        if world > 1:
            senv.gather()                       # (in fused mode the rows already went to the peers)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # keep the GPU busy long enough for the clock sampler to see it under load; every rank must run the SAME
    # number of extra steps (each step holds a cross-rank barrier / collective), so rank 0 decides
    remaining = max(0.0, 2.5 - (time.perf_counter() - t_wall0))
    rounds = torch.tensor([int(remaining / max(50 * (step_ms / K) * 1e-3, 1e-4)) + 1], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(rounds, src=0)
    for _ in range(min(int(rounds.item()), 2000)):
        run_steps(50, False, False)
    clocks = sampler.stop() if sampler else None

    # --- max over ranks
    t = torch.tensor([step_ms, kern_ms, warm_ms, warm_kern_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    ln = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    step_ms, kern_ms, warm_ms, warm_kern_ms, e2e_ms = t.tolist()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ms_per_step = step_ms / K
        kern_s = (kern_ms / K) * 1e-3
        events_kern_s, events_ms_per_step = kern_s, ms_per_step
        l2_text = "flushed between timed steps (256 MiB write, outside the event pairs)"
        if graph_ms is not None:
            ms_per_step = graph_ms / K
            kern_s = ms_per_step * 1e-3          # average launch duration over the timed region (incl. launch gaps)
            l2_text = ("cold by rotation: the K steps go round-robin over %d independent replicas of the batch "
                       "(%.1f MB each, %.0f MB > 2x the 126 MB L2); the K launches are one CUDA graph inside one event pair"
                       % (graph_info["replicas"], graph_info["replica_mb"],
                          graph_info["replicas"] * graph_info["replica_mb"]))
        achieved = E_local * bytes_per_env / kern_s / 1e9
        warm_kern_s = (warm_kern_ms / K) * 1e-3
        line = {
            "metric": "env-steps/s", "value": E_total / (ms_per_step * 1e-3), "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[wl], "envs_per_gpu": E_local, "envs_total": E_total,
                       "n_peds": cfg_global.n_peds, "n_samples": cfg_global.n_samples, "k_obstacles": cfg_global.k_obstacles,
                       "obs_dim": D, "parallelism": ("single GPU" if world == 1 else
                                                     "env-id sharding x%d, obs all-gather %s" % (world, {
                                                         "fused": "fused into the step kernel (bulk TMA stores into every peer's "
                                                                  "symmetric-memory buffer over NVLink); cross-rank barrier on a "
                                                                  "side stream, 3 rotating buffers",
                                                         "collective": "by one in-place ncclAllGather per step"}[gather_mode])),
                       "l2": l2_text,
                       "risk_block": ("faithful: the reference's own segmentation / tracker in float64, cn_faithful_kernel "
                                      "behind the step kernel (2 launches per step)" if args.risk_faithful else
                                      "intended: ideal association inside the step kernel"),
                       "actions": ("TD3 actor forward inside each step" if args.with_policy else
                                   "ring of 16 batches from a random-init TD3 actor + N(0,1) exploration noise, clipped")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(wl), "peak_source": peak_src,
                         "kernel": "%s<step>, %d worlds per CTA" % (env.kernel_name, env.kernel_tile), "kernel_us": kern_s * 1e6,
                         "algorithmic_bytes_per_env_step": bytes_per_env, "envs_per_launch": E_local,
                         "l2_warm": {"kernel_us": warm_kern_s * 1e6,
                                     "achieved": E_local * bytes_per_env / warm_kern_s / 1e9,
                                     "frac": E_local * bytes_per_env / warm_kern_s / 1e9 / peak}},
            "value_l2_warm": E_total / (warm_ms / K * 1e-3),
            "per_step_events": {"kernel_us": events_kern_s * 1e6, "ms_per_step": events_ms_per_step,
                                "value": E_total / (events_ms_per_step * 1e-3),
                                "frac": E_local * bytes_per_env / events_kern_s / 1e9 / peak,
                                "l2": "flushed between timed steps (256 MiB write, outside the event pairs); one event pair "
                                      "per step, which itself costs 2.6 us (profiles/tools/launch_overhead.py)"},
            "e2e": {"value": E_total / (e2e_ms * 1e-3 / K), "unit": "env-steps/s",
                    "h2d_bytes_per_step": env.h2d_bytes_per_step * world, "d2h_bytes_per_step": env.d2h_bytes_per_step * world,
                    "ms_per_step": e2e_ms / K, "api": "CrowdNavVecEnv.step_host (pinned host buffers)"},
            "gpu_launches": int(ln.item()),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl, args.risk_faithful)
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(wl: str, risk_faithful: bool = False):
    """The oracle port on the box's host cores: bounded sample of the same workload."""
    import numpy as np
    from crowdnav_b200.config import baseline_config
    from oracle.oracle import OracleEnv
    cores = os.cpu_count() or 1
    n = min(PER_GPU_ENVS[wl], 4096)
    cfg = baseline_config(WORKLOADS[wl], n_envs=n)
    if risk_faithful:
        cfg.flags |= 8
    env = OracleEnv(cfg, threads=cores)
    env.reset()
    rng = np.random.default_rng(0)
    acts = [np.stack([rng.uniform(0, 0.22, n), rng.uniform(-2, 2, n)], 1).astype(np.float32) for _ in range(8)]
    for i in range(3):
        env.step(acts[i])
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 10.0:
        env.step(acts[steps % 8])
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d steps x %d worlds (%.1f s), oracle/cn_oracle.c with OpenMP over worlds" % (steps, n, dt)}


if __name__ == "__main__":
    main()
