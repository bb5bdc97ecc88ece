#!/bin/bash
# Round 2 evidence for the shipped binary (1 GPU): the driver's own invocations, launch list, ncu --set full captures of
# the step kernel at c2 / c3 / c5 and of cn_faithful_kernel at c2, compute-sanitizer on the other two kernels, smoke().
mkdir -p gpurun_out
S=gpurun_out/r02_evidence_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step smoke 200 bash -c 'python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1'
step bench_driver 400 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_driver_k20.json 2> gpurun_out/r02_bench_c2_driver_k20.err'
step bench_ref 200 bash -c 'python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_reference_k20.json 2> gpurun_out/r02_bench_c2_reference_k20.err'
step bench_default 600 bash -c 'python bench.py > gpurun_out/r02_bench_c2_v12.json 2> gpurun_out/r02_bench_c2_v12.err'
step bench_faithful 300 bash -c 'python bench.py --risk-faithful --steps 100 --no-extras > gpurun_out/r02_bench_c2_v12_faithful.json 2> gpurun_out/r02_bench_c2_v12_faithful.err'
step launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_c2_v12.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_launches_c2_v12.log 2>&1
for spec in c2:529 c3:145 c5:65; do
  wl=${spec%%:*}; skip=${spec##*:}
  step ncu_full_$wl 300 ncu --set full --clock-control none --import-source on -k regex:cn_flat_kernel -s $skip -c 1 -f -o gpurun_out/r02_full_${wl}_v12 \
      python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_full_${wl}_v12.log 2>&1
done
step ncu_full_faithful 300 ncu --set full --clock-control none --import-source on -k regex:cn_faithful_kernel -s 529 -c 1 -f -o gpurun_out/r02_full_c2_faithful_v12 \
    python bench.py --risk-faithful --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_full_c2_faithful_v12.log 2>&1
for k in faithful warp flat; do for tool in memcheck racecheck; do
  step san_${tool}_$k 400 bash -c "compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_small.py $k > gpurun_out/r02_sanitizer_${tool}_${k}.log 2>&1"
done; done
cat $S
tail -n 3 gpurun_out/r02_smoke.log
for f in gpurun_out/r02_bench_c2_driver_k20 gpurun_out/r02_bench_c2_reference_k20 gpurun_out/r02_bench_c2_v12 gpurun_out/r02_bench_c2_v12_faithful; do echo "== $f"; head -c 6000 $f.json; echo; tail -n 5 $f.err; done
for k in faithful warp flat; do for t in memcheck racecheck; do echo "== $t $k"; tail -n 4 gpurun_out/r02_sanitizer_${t}_${k}.log; done; done
ls -la gpurun_out/*v12*.ncu-rep
