#!/bin/bash
# ray-group list capacity (per world: wall groups, pedestrian groups) and tile sweep: usage r02b_caps.sh <tag> "<wl>:<tile>:<caps> ..."   tile = auto | W,T ; caps = auto | gw,gp
tag=$1; shift
mkdir -p gpurun_out
for spec in $@; do
  IFS=':' read wl tile caps <<< "$spec"
  unset CN_FLAT_TILE CN_FLAT_CAPS
  [ "$tile" != auto ] && export CN_FLAT_TILE=$tile
  [ "$caps" != auto ] && export CN_FLAT_CAPS=$caps
  timeout 200 python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline --no-extras > gpurun_out/${tag}_${wl}_${tile}_${caps}.json 2> gpurun_out/${tag}_${wl}_${tile}_${caps}.err
done
python - "$tag" <<'PY'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/%s_*.json" % sys.argv[1])):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f" % (r["kernel_us"], r["frac"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
