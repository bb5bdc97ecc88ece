mkdir -p gpurun_out
S=gpurun_out/r01_run6_summary.txt; : > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step faithful_tests 300 python -m pytest tests/test_gpu_faithful.py -q > gpurun_out/pytest_gpu_faithful6.log 2>&1
step bench_c2_faithful 150 bash -c 'python bench.py --risk-faithful --steps 100 --no-cpu-baseline > gpurun_out/bench_c2_v12_faithful.json 2> gpurun_out/bench_c2_v12_faithful.err'
step full_faithful 150 ncu --set full --clock-control none --import-source on -k regex:cn_faithful_kernel -s 30 -c 1 -o gpurun_out/r01_full_faithful_v12 python bench.py --risk-faithful --timing events --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/full_faithful6.log 2>&1
cat $S; tail -3 gpurun_out/pytest_gpu_faithful6.log
