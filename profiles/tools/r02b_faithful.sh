#!/bin/bash
# risk_faithful kernel with shuffle scans / reductions: parity (kernel == oracle, bit for bit), smoke, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_faithful.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02b_pytest_gpu_faithful_v24.log; cat gpurun_out/r02b_pytest_gpu_faithful_v24.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 300 python bench.py --risk-faithful --steps 100 --no-extras > gpurun_out/r02b_bench_c2_v24_faithful.json 2> gpurun_out/r02b_bench_c2_v24_faithful.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02b_bench_c2_v24_faithful.json")); print("faithful: ms/step %.4f value %.4g e2e %.3g launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
PY
