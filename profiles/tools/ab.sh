#!/bin/bash
# A/B harness for the flat step kernel: parity tests, then bench.py over a list of CN_FLAT_TILE="W,threads" settings.
# usage (on the GPU box): bash profiles/tools/ab.sh <tag> "<cfg> <cfg> ..." "<workload> ..."   cfg = default | W,threads[:plain]
tag=${1:-ab}; cfgs=${2:-"default"}; wls=${3:-"c2 c3"}
mkdir -p gpurun_out
if [ -z "$CN_AB_NOTEST" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log; fi
for cfg in $cfgs; do
  for wl in $wls; do
    tile=${cfg%%:*}; store=tma; [[ "$cfg" == *:* ]] && store=${cfg##*:}
    if [ "$tile" = "default" ]; then unset CN_FLAT_TILE; else export CN_FLAT_TILE=$tile; fi
    export CN_FLAT_STORE=$store
    python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/${tag}_bench_${wl}_${cfg}.json 2> gpurun_out/${tag}_bench_${wl}_${cfg}.err
  done
done
python - "$tag" <<'PY'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json" % sys.argv[1])):
    try:
        d=json.load(open(f)); print(f, d["roofline"]["kernel"], "graph", round(d["roofline"]["kernel_us"],2), round(d["roofline"]["frac"],4), "events", round(d.get("per_step_events",{}).get("kernel_us",0),2))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
