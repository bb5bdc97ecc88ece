#!/bin/bash
# Final validation of a round on one GPU box: what the driver runs (gpu tests, smoke, bench, reference arm) + c3 / c5 lines.
mkdir -p gpurun_out
S=gpurun_out/r01_validate_summary.txt; : > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step gpu_tests 420 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1
step smoke 120 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_final.log 2>&1
step bench_c2 240 bash -c 'python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err'
step bench_ref 240 bash -c 'python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_c2_reference_final.json 2> gpurun_out/bench_c2_reference_final.err'
step bench_c5 150 bash -c 'python bench.py --workload c5 --no-cpu-baseline > gpurun_out/bench_c5_final.json 2> gpurun_out/bench_c5_final.err'
step bench_c3 150 bash -c 'python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err'
cat $S; tail -3 gpurun_out/pytest_gpu_final.log; tail -2 gpurun_out/smoke_final.log
