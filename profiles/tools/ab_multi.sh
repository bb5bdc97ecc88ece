#!/bin/bash
# multi-GPU A/B of the fused-gather step: bash profiles/tools/ab_multi.sh <tag> <n_gpus> "<cfg> ..."   cfg = default | W,threads[:plain]
tag=$1; n=$2; cfgs=${3:-default}
mkdir -p gpurun_out
port=29520
for cfg in $cfgs; do
  tile=${cfg%%:*}; store=tma; [[ "$cfg" == *:* ]] && store=${cfg##*:}
  if [ "$tile" = "default" ]; then unset CN_FLAT_TILE; else export CN_FLAT_TILE=$tile; fi
  export CN_FLAT_STORE=$store
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 50 --warmup 5 \
      > gpurun_out/${tag}_bench_c2_${n}gpu_${cfg}.json 2> gpurun_out/${tag}_bench_c2_${n}gpu_${cfg}.err
done
python - "$tag" <<'PY'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/%s_bench_*gpu_*.json" % sys.argv[1])):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "ms/step %.4f kernel_us %.2f value %.4g" % (d["ms_per_step"], r["kernel_us"], d["value"]))
    except Exception as e: print(f, "ERR", e)
PY
