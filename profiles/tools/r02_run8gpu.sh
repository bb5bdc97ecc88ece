#!/bin/bash
# 8 GPUs: the driver's N=8 invocation (fused_async gather, self-check, configs[3] as an extra key), the end-of-kernel
# variant for comparison, then N=4 on the same box.
mkdir -p gpurun_out
S=gpurun_out/r02_run8gpu_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
nvidia-smi -L | wc -l >> $S
tr() { # n port args...
  local n=$1 port=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"
}
export -f tr
step bench_8gpu 420 bash -c "tr 8 29501 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_8gpu_async.json 2> gpurun_out/r02_bench_c2_8gpu_async.err"
step bench_8gpu_fused 300 bash -c "tr 8 29502 --steps 20 --warmup 5 --gather fused --no-extras > gpurun_out/r02_bench_c2_8gpu_fused.json 2> gpurun_out/r02_bench_c2_8gpu_fused.err"
step bench_4gpu 300 bash -c "tr 4 29503 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_4gpu_async.json 2> gpurun_out/r02_bench_c2_4gpu_async.err"
step bench_4gpu_fused 300 bash -c "tr 4 29504 --steps 20 --warmup 5 --gather fused --no-extras > gpurun_out/r02_bench_c2_4gpu_fused.json 2> gpurun_out/r02_bench_c2_4gpu_fused.err"
step bench_2gpu 300 bash -c "tr 2 29505 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_2gpu_async.json 2> gpurun_out/r02_bench_c2_2gpu_async.err"
cat $S
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_c2_?gpu_*.json")):
    try:
        d=json.load(open(f))
        print(f, "N=%d ms/step %.4f value %.4g replicas_only %.4f verified %s e2e %.3g" % (d["n_gpus"], d["ms_per_step"], d["value"], d.get("replicas_only",{}).get("ms_per_step",0), d.get("gather_verified"), d["e2e"]["value"]), d.get("gather_fallback"), d.get("gather_check"))
        if "configs3" in d: print("   configs3:", json.dumps(d["configs3"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
