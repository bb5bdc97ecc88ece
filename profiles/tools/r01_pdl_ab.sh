#!/bin/bash
# Programmatic dependent launch of the step kernel: parity tests with it on, bench A/B (CN_PDL=0 turns it off).
mkdir -p gpurun_out
S=gpurun_out/r01_pdl_summary.txt; : > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step gpu_tests 420 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_pdl.log 2>&1
step bench_c2_pdl 200 bash -c 'python bench.py --no-cpu-baseline > gpurun_out/bench_c2_v8_pdl.json 2> gpurun_out/bench_c2_v8_pdl.err'
step bench_c2_nopdl 200 bash -c 'CN_PDL=0 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_v8_nopdl.json 2> gpurun_out/bench_c2_v8_nopdl.err'
step bench_c3_pdl 200 bash -c 'python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_v8_pdl.json 2> gpurun_out/bench_c3_v8_pdl.err'
cat $S; tail -3 gpurun_out/pytest_gpu_pdl.log
python - <<'PY'
import json
for f in ["bench_c2_v8_pdl","bench_c2_v8_nopdl","bench_c3_v8_pdl"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f,"ms/step %.4f"%d["ms_per_step"],"events_us %.2f"%d["per_step_events"]["kernel_us"],"warm_us %.2f"%d["roofline"]["l2_warm"]["kernel_us"],"e2e %.3e"%d["e2e"]["value"])
    except Exception as e: print(f,"ERR",e)
PY
