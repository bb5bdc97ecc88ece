#!/bin/bash
# 2 GPUs: the 16-bit wire format -- gather parity (bit patterns) in every mode, bench async16 vs async
mkdir -p gpurun_out
S=gpurun_out/r02_run9_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step multi_tests 500 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_pytest_multi_d.log 2>&1
port=30100
for mode in fused_async16 fused_async; do
  port=$((port+1))
  step bench_2gpu_$mode 300 bash -c "python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 60 --warmup 5 --gather $mode --no-extras > gpurun_out/r02_bench_c2_2gpu_${mode}_d.json 2> gpurun_out/r02_bench_c2_2gpu_${mode}_d.err"
done
port=$((port+1))
step bench_2gpu_c4 300 bash -c "python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 --workload c4 --no-extras > gpurun_out/r02_bench_c4_2gpu_d.json 2> gpurun_out/r02_bench_c4_2gpu_d.err"
cat $S
tail -n 25 gpurun_out/r02_pytest_multi_d.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*_d.json")):
    try:
        d=json.load(open(f))
        print(f, "ms/step %.4f value %.4g replicas_only %.4f verified %s e2e %.3g ev_ms %.4f launches %d" % (d["ms_per_step"], d["value"], d.get("replicas_only",{}).get("ms_per_step",0), d.get("gather_verified"), d["e2e"]["value"], d["per_step_events"]["ms_per_step"], d["gpu_launches"]), d.get("gather_fallback"), d.get("gather_check"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
