#!/bin/bash
# 1 GPU: parity (incl. the robot contact model), then A/B of resident CTAs per SM (register cap + tile choice): 4 / 5 / 6
mkdir -p gpurun_out
S=gpurun_out/r02_run8_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step gpu_tests 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_e.log 2>&1
for lib in libcrowdnav.so libcrowdnav_mb5.so libcrowdnav_mb6.so; do
  for wl in c2 c3 c5; do
    step bench_${wl}_$lib 240 bash -c "CN_LIB=$lib python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline --no-extras > gpurun_out/r02_bench_${wl}_v11_$lib.json 2> gpurun_out/r02_bench_${wl}_v11_$lib.err"
  done
done
step timeline_c2 200 bash -c 'python profiles/tools/timeline_flat.py c2 > gpurun_out/r02_timeline_c2_v11.txt 2>&1'
cat $S
tail -n 15 gpurun_out/r02_pytest_gpu_e.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_c*_v11_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f | events %.2f | step_n %.2f" % (r["kernel_us"], r["frac"], d["per_step_events"]["kernel_us"], r["l2_warm"]["kernel_us"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
head -24 gpurun_out/r02_timeline_c2_v11.txt
