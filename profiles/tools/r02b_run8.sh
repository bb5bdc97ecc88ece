#!/bin/bash
# v18 validation: the whole GPU suite, smoke(), c5 tile check, automatic-tile bench lines at c2 / c3 / c5
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu_v18.log 2>&1; tail -n 4 gpurun_out/r02b_pytest_gpu_v18.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke_v18.log 2>&1; tail -n 2 gpurun_out/r02b_smoke_v18.log
bash profiles/tools/r02b_caps.sh r02b_v18 c2:auto:auto c3:auto:auto c5:auto:auto c5:12,384:auto c5:13,384:auto c5:8,256:auto
