#!/bin/bash
# parity first (flat-kernel test files), then same-box A/B of library builds
tag=$1; libs=$2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi2.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log
bash profiles/tools/r02b_libab.sh $tag "$libs"
