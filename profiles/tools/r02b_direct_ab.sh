#!/bin/bash
# Round 2 (second session): the "direct rows" step kernel (rows straight to global memory, no staging tile).
# Parity first (tile shapes incl. odd tiles / 384- and 512-thread CTAs, unaligned row block, the staged instance via
# CN_FLAT_DIRECT=0), then bench.py over tiles: cfg = d0 (staged, automatic tile) | auto (direct, automatic) | W,threads[:plain]
# usage (GPU box): bash profiles/tools/r02b_direct_ab.sh <tag> "<wl>=<cfg>,<cfg>;..." e.g. "c2=d0 6,256;c3=d0 12,256 28,384"
tag=${1:-r02b_ab}; spec=${2:-"c2=d0 auto"}
mkdir -p gpurun_out
if [ -z "$CN_AB_NOTEST" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi2.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log
fi
IFS=';' read -ra parts <<< "$spec"
for part in "${parts[@]}"; do
  wl=${part%%=*}; cfgs=${part#*=}
  for cfg in $cfgs; do
    unset CN_FLAT_TILE CN_FLAT_DIRECT CN_FLAT_STORE
    if [ "$cfg" = "d0" ]; then export CN_FLAT_DIRECT=0
    elif [ "$cfg" != "auto" ]; then
      tile=${cfg%%:*}; export CN_FLAT_TILE=$tile
      [[ "$cfg" == *:plain ]] && export CN_FLAT_STORE=plain
    fi
    timeout 200 python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_${wl}_${cfg}.json 2> gpurun_out/${tag}_bench_${wl}_${cfg}.err
  done
done
python - "$tag" <<'PY'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json" % sys.argv[1])):
    try:
        d=json.load(open(f)); print(f, d["roofline"]["kernel"], "graph us", round(d["roofline"]["kernel_us"],2), "frac", round(d["roofline"]["frac"],4), "events", round(d.get("per_step_events",{}).get("kernel_us",0),2))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
