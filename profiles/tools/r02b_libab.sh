#!/bin/bash
# A/B of library builds (CN_LIB) on bench.py: usage r02b_libab.sh <tag> "<lib> <lib> ..." "<wl> ..."
tag=$1; libs=$2; wls=${3:-"c2 c3 c5"}
mkdir -p gpurun_out
for lib in $libs; do for wl in $wls; do
  CN_LIB=$lib timeout 200 python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline --no-extras > gpurun_out/${tag}_${wl}_${lib}.json 2> gpurun_out/${tag}_${wl}_${lib}.err
done; done
python - "$tag" <<'PY'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/%s_*.json" % sys.argv[1])):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f | events %.2f" % (r["kernel_us"], r["frac"], d["per_step_events"]["kernel_us"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
