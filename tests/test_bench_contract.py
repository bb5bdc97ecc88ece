"""The bench.py contract, as far as it can be checked without a GPU: the reference arm runs here (oracle port on the
host cores) and prints ONE JSON line with the agreed keys; the committed B200 line of the round carries `roofline`,
`cpu_baseline`, `e2e`, `clocks` and `gpu_launches`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/s" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in d["config"]["workload"]


def test_committed_b200_line_has_the_contract_keys():
    d = json.loads(open(os.path.join(ROOT, "profiles", "r01", "bench_c2_v7.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["algorithmic_bytes_per_env_step"] == 4 * (2 * 16 + 2 * 8 * 20 + 2 + (359 + 7 + 32) + 1) + 1 == 3013
    assert abs(r["achieved"] - 4096 * 3013 / (r["kernel_us"] * 1e-6) / 1e9) < 1e-6 * r["achieved"]
    assert d["gpu_launches"] == d["steps"] and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["e2e"]["h2d_bytes_per_step"] == 4096 * 8 and d["e2e"]["d2h_bytes_per_step"] == 4096 * (398 * 4 + 5)
    assert d["e2e"]["value"] < d["value"] and not d["clocks"]["reasons"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
