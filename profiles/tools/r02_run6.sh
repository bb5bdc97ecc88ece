#!/bin/bash
# 1 GPU: parity of the strip prefilter / bulk fill, then bench c2 / c3 and a tile-shape sweep.
mkdir -p gpurun_out
S=gpurun_out/r02_run6_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step gpu_tests 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c.log 2>&1
run() { # workload tile
  if [ "$2" = default ]; then unset CN_FLAT_TILE; else export CN_FLAT_TILE=$2; fi
  local t0=$(date +%s)
  timeout 200 python bench.py --steps 100 --warmup 5 --workload $1 --no-cpu-baseline --no-extras > gpurun_out/r02_tile_$1_$2.json 2> gpurun_out/r02_tile_$1_$2.err
  echo "tile_$1_$2 rc=$? $(( $(date +%s) - t0 ))s" >> $S
}
for t in default 4,128 6,192 10,256 12,384 16,512; do run c2 $t; done
for t in default 10,256 16,256 20,384 28,512 32,512 8,128; do run c3 $t; done
unset CN_FLAT_TILE
cat $S
tail -n 15 gpurun_out/r02_pytest_gpu_c.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_tile_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f | events %.2f | step_n %.2f" % (r["kernel_us"], r["frac"], d["per_step_events"]["kernel_us"], r["l2_warm"]["kernel_us"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
