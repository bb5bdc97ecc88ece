#!/bin/bash
# 1 GPU: all GPU tests, ncu --set full of the step kernel at c2 / c3 (cold-rotation launch), tile-shape sweep.
mkdir -p gpurun_out
S=gpurun_out/r02_run5_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step gpu_tests 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_b.log 2>&1
for wl in c2 c3; do
  skip=$([ $wl = c2 ] && echo 529 || echo 145)   # 16 warm-start steps + (R-1) x 16 replica steps, then the cold rotation
  step ncu_full_$wl 300 ncu --set full --clock-control none --import-source on -k regex:cn_flat_kernel -s $skip -c 1 -f -o gpurun_out/r02_full_${wl}_v8 \
      python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_full_${wl}_v8.log 2>&1
done
run() { # tag workload tile
  if [ "$3" = default ]; then unset CN_FLAT_TILE; else export CN_FLAT_TILE=$3; fi
  timeout 200 python bench.py --steps 100 --warmup 5 --workload $2 --no-cpu-baseline --no-extras > gpurun_out/r02_tile_$2_$3.json 2> gpurun_out/r02_tile_$2_$3.err
}
for t in default 4,128 6,192 10,256 12,384 16,512; do step tile_c2_$t 220 run x c2 $t; done
for t in default 10,256 16,256 20,384 28,512 32,512 8,128; do step tile_c3_$t 220 run x c3 $t; done
unset CN_FLAT_TILE
cat $S
tail -n 4 gpurun_out/r02_pytest_gpu_b.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_tile_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f | events %.2f | step_n %.2f" % (r["kernel_us"], r["frac"], d["per_step_events"]["kernel_us"], r["l2_warm"]["kernel_us"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
ls -la gpurun_out/*.ncu-rep | tail -3
