#!/bin/bash
# direct rows with the fill moved off the critical path (pose warp 0, before #A): parity, small tile sweep, timeline
export CN_AB_NOTEST=
bash profiles/tools/r02b_direct_ab.sh r02b_ab4 "c2=6,256 8,256;c3=19,256 28,384;c5=8,256 12,384"
CN_FLAT_TILE=28,384 timeout 200 python profiles/tools/timeline_flat.py c3 > gpurun_out/r02b_timeline4_c3_28,384.txt 2>&1
tail -n 18 gpurun_out/r02b_timeline4_c3_28,384.txt
