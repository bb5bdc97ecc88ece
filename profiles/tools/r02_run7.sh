#!/bin/bash
# 1 GPU: parity after the latency work (inline E-J, group-reduced atomics, (world, face) lanes), bench, timeline, sanitizer.
mkdir -p gpurun_out
S=gpurun_out/r02_run7_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step gpu_tests 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_d.log 2>&1
for wl in c2 c3 c5; do
  step bench_$wl 240 bash -c "python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline --no-extras > gpurun_out/r02_bench_${wl}_v10.json 2> gpurun_out/r02_bench_${wl}_v10.err"
done
step timeline_c2 200 bash -c 'python profiles/tools/timeline_flat.py c2 > gpurun_out/r02_timeline_c2_v10.txt 2>&1'
step timeline_c3 200 bash -c 'python profiles/tools/timeline_flat.py c3 > gpurun_out/r02_timeline_c3_v10.txt 2>&1'
for tool in memcheck racecheck synccheck; do
  step san_${tool}_flat 400 bash -c "compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_small.py flat > gpurun_out/r02_sanitizer_${tool}_flat.log 2>&1"
done
cat $S
tail -n 15 gpurun_out/r02_pytest_gpu_d.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_c*_v10.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f, r["kernel"], "graph %.2f us frac %.4f | events %.2f | step_n %.2f | e2e %s" % (r["kernel_us"], r["frac"], d["per_step_events"]["kernel_us"], r["l2_warm"]["kernel_us"], {k: round(v["ms_per_step"]*1e3,1) for k,v in d["e2e"]["modes"].items()}))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
head -24 gpurun_out/r02_timeline_c2_v10.txt
for t in memcheck racecheck synccheck; do echo "== $t"; tail -n 6 gpurun_out/r02_sanitizer_${t}_flat.log; done
