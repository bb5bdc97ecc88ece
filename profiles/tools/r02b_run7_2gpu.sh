#!/bin/bash
# 2 GPUs, v19: gather parity in every mode (staged instance, against the direct-rows instance on one GPU), bench lines
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02b_pytest_multi_2gpu_v19.log 2>&1; tail -n 6 gpurun_out/r02b_pytest_multi_2gpu_v19.log
port=30200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 60 --warmup 5 --no-extras > gpurun_out/r02b_bench_c2_2gpu_v19.json 2> gpurun_out/r02b_bench_c2_2gpu_v19.err
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 --workload c4 --no-extras > gpurun_out/r02b_bench_c4_2gpu_v19.json 2> gpurun_out/r02b_bench_c4_2gpu_v19.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02b_bench_*2gpu_v19.json")):
    try:
        d=json.load(open(f))
        print(f, "ms/step %.4f value %.4g replicas_only %.4f verified %s e2e %.3g launches %d" % (d["ms_per_step"], d["value"], d.get("replicas_only",{}).get("ms_per_step",0), d.get("gather_verified"), d["e2e"]["value"], d["gpu_launches"]), d.get("gather_fallback"), d.get("gather_check"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
