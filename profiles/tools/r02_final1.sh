#!/bin/bash
# Round 2, final single-GPU validation of the shipped tree: all GPU tests, smoke, the driver's bench invocations,
# launch list of the timed region, sanitizer on the step kernel.
mkdir -p gpurun_out
S=gpurun_out/r02_final1_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
step gpu_tests 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1
step smoke 200 bash -c 'python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1'
step bench_ref 200 bash -c 'python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_reference_final.json 2> gpurun_out/r02_bench_c2_reference_final.err'
step bench_driver 400 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_final_k20.json 2> gpurun_out/r02_bench_c2_final_k20.err'
step bench_default 600 bash -c 'python bench.py > gpurun_out/r02_bench_c2_final.json 2> gpurun_out/r02_bench_c2_final.err'
step launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 740 -c 120 --csv --log-file gpurun_out/r02_launches_c2_final.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_launches_c2_final.log 2>&1
for tool in memcheck racecheck; do
  step san_${tool}_flat 400 bash -c "compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_small.py flat > gpurun_out/r02_sanitizer_${tool}_flat_final.log 2>&1"
  step san_${tool}_warp 400 bash -c "compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_small.py warp > gpurun_out/r02_sanitizer_${tool}_warp_final.log 2>&1"
done
cat $S
tail -n 12 gpurun_out/r02_pytest_gpu_final.log
tail -n 2 gpurun_out/r02_smoke_final.log
python - <<'PY'
import json
for f in ("r02_bench_c2_reference_final","r02_bench_c2_final_k20","r02_bench_c2_final"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print("==",f,"value %.4g ms/step %.5f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
        for k in ("roofline_c3","rollout_td3","value_l2_warm","value_l2_warm_lib_graph"):
            if k in d: print("   ",k,json.dumps(d[k])[:700])
    except Exception as e:
        print(f,"ERR",e); print(open("gpurun_out/%s.err"%f).read()[-2000:])
PY
python profiles/tools/summarise_launches.py gpurun_out/r02_launches_c2_final.csv "timed region of python bench.py --steps 20 --warmup 3 (launches 740..860)" | head -12
for k in flat warp; do for t in memcheck racecheck; do echo "== $t $k"; tail -n 2 gpurun_out/r02_sanitizer_${t}_${k}_final.log; done; done
