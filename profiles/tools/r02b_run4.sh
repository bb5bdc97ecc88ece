#!/bin/bash
# usage: r02b_run4.sh <tag>   -- parity (flat-kernel files), bench c2/c3/c5 with the automatic tiles, timeline at c3
tag=${1:-r02b_v15}
mkdir -p gpurun_out
bash profiles/tools/r02b_direct_ab.sh $tag "c2=auto;c3=auto;c5=auto"
timeout 200 python profiles/tools/timeline_flat.py c3 > gpurun_out/${tag}_timeline_c3.txt 2>&1
sed -n 6,17p gpurun_out/${tag}_timeline_c3.txt; tail -n 18 gpurun_out/${tag}_timeline_c3.txt
