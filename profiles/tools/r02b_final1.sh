#!/bin/bash
# Final evidence for the shipped binary v19 (1 GPU): whole GPU suite, smoke(), the driver's own invocations (our arm + reference arm),
# the default bench line, c3 / c5 / risk_faithful lines, compute-sanitizer on both instances of the flat kernel, ncu --set full at c2 / c3 / c5
mkdir -p gpurun_out
S=gpurun_out/r02b_final_summary.txt; : > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step gpu_tests 900 bash -c 'python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu_v19.log 2>&1'
step smoke 300 bash -c 'python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke_v19.log 2>&1'
step bench_ref 200 bash -c 'python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02b_bench_c2_v19_reference_arm_k20.json 2> gpurun_out/r02b_bench_c2_v19_reference_arm_k20.err'
step bench_driver 400 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02b_bench_c2_v19_driver_k20.json 2> gpurun_out/r02b_bench_c2_v19_driver_k20.err'
step bench_default 600 bash -c 'python bench.py > gpurun_out/r02b_bench_c2_v19.json 2> gpurun_out/r02b_bench_c2_v19.err'
for wl in c3 c5; do
  step bench_$wl 300 bash -c "python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline --no-extras > gpurun_out/r02b_bench_${wl}_v19.json 2> gpurun_out/r02b_bench_${wl}_v19.err"
done
step bench_faithful 300 bash -c 'python bench.py --risk-faithful --steps 100 --no-extras > gpurun_out/r02b_bench_c2_v19_faithful.json 2> gpurun_out/r02b_bench_c2_v19_faithful.err'
for mode in direct staged; do for tool in memcheck racecheck; do
  if [ $mode = staged ]; then export CN_FLAT_DIRECT=0; else unset CN_FLAT_DIRECT; fi
  step san_${tool}_$mode 300 bash -c "compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_small.py flat > gpurun_out/r02b_sanitizer_${tool}_flat_${mode}_v19.log 2>&1"
done; done
unset CN_FLAT_DIRECT
for spec in c2:529 c3:145 c5:65; do
  wl=${spec%%:*}; skip=${spec##*:}
  step ncu_full_$wl 300 ncu --set full --clock-control none --import-source on -k regex:cn_flat_kernel -s $skip -c 1 -f -o gpurun_out/r02b_v19_full_${wl} \
      python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02b_v19_full_${wl}.log 2>&1
done
cat $S
tail -n 3 gpurun_out/r02b_pytest_gpu_v19.log; tail -n 2 gpurun_out/r02b_smoke_v19.log
for f in gpurun_out/r02b_bench_c2_v19_driver_k20 gpurun_out/r02b_bench_c2_v19_reference_arm_k20 gpurun_out/r02b_bench_c2_v19; do echo "== $f"; head -c 5000 $f.json; echo; tail -n 3 $f.err; done
for mode in direct staged; do for t in memcheck racecheck; do echo "== $t $mode"; tail -n 1 gpurun_out/r02b_sanitizer_${t}_flat_${mode}_v19.log; done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02b_bench_c[35]_v19.json"))+["gpurun_out/r02b_bench_c2_v19_faithful.json"]:
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f" % (r["kernel_us"], r["frac"]), "value %.4g" % d["value"])
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
