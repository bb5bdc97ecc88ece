#!/bin/bash
# v16 (32 strips, aligned pose quad, fill tile from the leftover budget): parity, same-box A/B against v14, c3 one-wave tiles, sanitizers
tag=r02b_v16
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi2.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log
bash profiles/tools/r02b_libab.sh ${tag}_ab "libcrowdnav_v14.so libcrowdnav.so"
CN_AB_NOTEST=1 bash profiles/tools/r02b_direct_ab.sh ${tag}_tiles "c3=28,384 19,256"
for mode in direct staged; do for tool in memcheck racecheck; do
  if [ $mode = staged ]; then export CN_FLAT_DIRECT=0; else unset CN_FLAT_DIRECT; fi
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_small.py flat > gpurun_out/${tag}_sanitizer_${tool}_flat_${mode}.log 2>&1
  echo "== $tool $mode"; tail -n 2 gpurun_out/${tag}_sanitizer_${tool}_flat_${mode}.log
done; done
