#!/bin/bash
# 8 GPUs: 16-bit wire format at N = 8 (driver invocation, configs[3] extra) and N = 4
mkdir -p gpurun_out
S=gpurun_out/r02_run10_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
tr() { local n=$1 port=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
export -f tr
step bench_8gpu_16 420 bash -c "tr 8 29601 --steps 20 --warmup 5 --gather fused_async16 > gpurun_out/r02_bench_c2_8gpu_async16.json 2> gpurun_out/r02_bench_c2_8gpu_async16.err"
step bench_4gpu_16 300 bash -c "tr 4 29602 --steps 20 --warmup 5 --gather fused_async16 --no-extras > gpurun_out/r02_bench_c2_4gpu_async16.json 2> gpurun_out/r02_bench_c2_4gpu_async16.err"
cat $S
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_c2_?gpu_async16.json")):
    try:
        d=json.load(open(f))
        print(f, "N=%d ms/step %.4f value %.4g replicas_only %.4f verified %s e2e %.3g launches %d" % (d["n_gpus"], d["ms_per_step"], d["value"], d.get("replicas_only",{}).get("ms_per_step",0), d.get("gather_verified"), d["e2e"]["value"], d["gpu_launches"]), d.get("gather_fallback"), d.get("gather_check"))
        if "configs3" in d: print("   configs3:", json.dumps(d["configs3"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
