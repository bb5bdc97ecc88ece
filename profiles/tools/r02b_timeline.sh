#!/bin/bash
# %globaltimer phase timeline of the direct-rows step kernel at c3 for the two one-wave tiles (needs libcrowdnav_timeline.so)
mkdir -p gpurun_out
for cfg in "c3 28,384" "c3 19,256" "c2 8,256"; do
  set -- $cfg
  CN_FLAT_TILE=$2 timeout 200 python profiles/tools/timeline_flat.py $1 > gpurun_out/r02b_timeline_$1_$2.txt 2>&1
  echo "== $cfg"; tail -n 32 gpurun_out/r02b_timeline_$1_$2.txt
done
