#!/bin/bash
# diagnostics: where does the fused gather's time go?  CN_GATHER_DEBUG bits: 1 no guard, 2 no signal, 4 no data
mkdir -p gpurun_out
port=29800
for mode in fused_async fused; do
for dbg in 0 1 2 3 4 7; do
  port=$((port+1))
  CN_GATHER_DEBUG=$dbg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 60 --warmup 5 --gather $mode --no-extras --skip-verify > gpurun_out/r02_dbg_${mode}_$dbg.json 2> gpurun_out/r02_dbg_${mode}_$dbg.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_dbg_*.json")):
    try:
        d=json.load(open(f)); print(f, "graph ms/step %.4f  events kernel_us %.2f step %.4f  replicas_only %.4f" % (d["ms_per_step"], d["per_step_events"]["kernel_us"], d["per_step_events"]["ms_per_step"], d["replicas_only"]["ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
