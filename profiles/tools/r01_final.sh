#!/bin/bash
# One GPU-box call at the end of round 1: parity of the risk_faithful kernel, bench lines of the final kernels,
# launch lists.  Every step is bounded by its own timeout and writes under gpurun_out/.
mkdir -p gpurun_out
S=gpurun_out/r01_final_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step faithful_tests 300 python -m pytest tests/test_gpu_faithful.py -q > gpurun_out/pytest_gpu_faithful.log 2>&1
step bench_c2 240 bash -c 'python bench.py > gpurun_out/bench_c2_v7.json 2> gpurun_out/bench_c2_v7.err'
step bench_c2_faithful 150 bash -c 'python bench.py --risk-faithful --steps 100 > gpurun_out/bench_c2_v7_faithful.json 2> gpurun_out/bench_c2_v7_faithful.err'
step gpu_tests 420 python -m pytest tests -m gpu -q --deselect tests/test_gpu_faithful.py > gpurun_out/pytest_gpu.log 2>&1
step launches_faithful 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_c2_v7_faithful.csv python bench.py --risk-faithful --timing events --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches_c2_v7_faithful.log 2>&1
step launches_c2 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_c2_v7.csv python bench.py --timing events --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches_c2_v7.log 2>&1
step smoke 120 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1
step bench_c3 150 bash -c 'python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_v7.json 2> gpurun_out/bench_c3_v7.err'
step full_faithful 150 ncu --set full --clock-control none --import-source on -k regex:cn_faithful_kernel -s 30 -c 1 \
    -o gpurun_out/r01_full_faithful python bench.py --risk-faithful --timing events --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/full_faithful.log 2>&1
cat $S
tail -3 gpurun_out/pytest_gpu_faithful.log gpurun_out/pytest_gpu.log
