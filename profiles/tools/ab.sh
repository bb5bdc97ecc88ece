python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/s5_pytest.log; cat gpurun_out/s5_pytest.log
for cfg in "8,256" "6,256" "16,512" "14,512"; do
  for wl in c2 c3; do
    CN_FLAT_TILE=$cfg python bench.py --steps 40 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/s5_bench_${wl}_${cfg}.json 2> gpurun_out/s5_bench_${wl}_${cfg}.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s5_bench_*.json")):
    try:
        d=json.load(open(f)); print(f, d["roofline"]["kernel"], round(d["roofline"]["kernel_us"],2), round(d["roofline"]["frac"],4), round(d["roofline"]["l2_warm"]["kernel_us"],2))
    except Exception as e: print(f, "ERR", e)
PY
