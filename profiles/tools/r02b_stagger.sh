#!/bin/bash
# de-phasing experiment: CN_FLAT_STAGGER_NS sweep    usage: r02b_stagger.sh <tag> "<wl>:<ns> ..."
tag=$1; shift
mkdir -p gpurun_out
for spec in $@; do
  IFS=':' read wl ns <<< "$spec"
  CN_FLAT_STAGGER_NS=$ns timeout 200 python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline --no-extras > gpurun_out/${tag}_${wl}_${ns}.json 2> gpurun_out/${tag}_${wl}_${ns}.err
done
python - "$tag" <<'PY'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/%s_*.json" % sys.argv[1])):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f" % (r["kernel_us"], r["frac"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
