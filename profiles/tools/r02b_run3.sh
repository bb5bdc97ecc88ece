#!/bin/bash
# direct rows (TMA fill), automatic tiles: the whole GPU test suite, default bench lines, ncu --set full (cold + warm caches) at c3 and c2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu_v14.log 2>&1; tail -n 4 gpurun_out/r02b_pytest_gpu_v14.log
for wl in c2 c3 c5; do
  timeout 300 python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline --no-extras > gpurun_out/r02b_bench_${wl}_v14.json 2> gpurun_out/r02b_bench_${wl}_v14.err
done
for spec in c3:145 c2:529; do
  wl=${spec%%:*}; skip=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:cn_flat_kernel -s $skip -c 1 -f -o gpurun_out/r02b_full_${wl}_v14 \
      python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02b_full_${wl}_v14.log 2>&1
  timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:cn_flat_kernel -s $skip -c 1 -f -o gpurun_out/r02b_warm_${wl}_v14 \
      python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02b_warm_${wl}_v14.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02b_bench_c*_v14.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f | events %.2f" % (r["kernel_us"], r["frac"], d["per_step_events"]["kernel_us"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
ls -la gpurun_out/*.ncu-rep
