"""Fixture for the `risk_faithful` block: replays the recorded physics of the reference-in-the-loop traces
(tests/golden/trace_*.npz: odometry + raw scans) through the REFERENCE's Env once more and records what the
traces do not hold: the safety counters (ENV:653-654, 998-1005) and the size of its tracker dict after every
get_state.  Run in the build container only (needs /root/reference):

    python tests/gen_golden_faithful.py        ->  tests/golden/faithful_counters.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from ref_harness import DEFAULT_PARAMS, Reference, Scan      # noqa: E402
from trace_configs import TRACES, trace_config   # noqa: E402


def replay(ref, name):
    cfg, params, _ = trace_config(name)
    ref.params.clear()
    ref.params.update(DEFAULT_PARAMS)       # every trace starts from the YAML defaults
    ref.params.update(params)
    z = np.load(os.path.join(HERE, "golden", "trace_%s.npz" % name))
    n = len(z["odom"])
    holder = {}
    ref.scan_source = lambda: Scan(holder["scan"])
    out = np.zeros((n, 4), dtype=np.int32)       # ego, social, obstacle-present, len(tracked_obstacles)
    env, step = None, 0
    for t in range(n):
        if z["episode_start"][t] > 0:
            env = ref.make_env(max_step=cfg.max_steps, k_obstacle_count=cfg.k_obstacles)
            ref.set_odom(env, *[float(q) for q in z["odom"][t]])
            holder["scan"] = [float(q) for q in z["scan"][t]]
            s = env.reset()
            env.done = False
            step = 0
        else:
            def physics(dt, t=t):
                ref.set_odom(env, *[float(q) for q in z["odom"][t]])
                holder["scan"] = [float(q) for q in z["scan"][t]]
            ref.clock.on_sleep = physics
            s, r, d = env.step([float(z["action"][t][0]), float(z["action"][t][1])], step + 1, mode="continuous")
            ref.clock.on_sleep = None
            step += 1
            assert float(r) == float(z["ref_reward"][t]) and float(d) == float(z["ref_done"][t])
        assert np.array_equal(np.asarray(s, dtype=np.float64), z["ref_state"][t]), "replay diverged from the trace"
        out[t] = (env.ego_safety_violation_count, env.social_safety_violation_count,
                  env.obstacle_present_step_counts, len(env.tracked_obstacles))
    return out


def main():
    ref = Reference()
    arrays = {}
    for name in TRACES:
        arrays[name] = replay(ref, name)
        print(name, "rows", len(arrays[name]), "final-row maxima", arrays[name].max(0))
    np.savez_compressed(os.path.join(HERE, "golden", "faithful_counters.npz"), **arrays)


if __name__ == "__main__":
    main()
