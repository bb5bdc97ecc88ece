#!/bin/bash
# ncu evidence for the shipped binary: launch list of a driver-style bench run + --set full of the step kernel at c2 / c3 / c5
# (cold: ncu flushes all caches before each pass; warm: --cache-control none)   usage: r02b_capture.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_launches_c2.log 2>&1
for spec in c2:529 c3:145 c5:65; do
  wl=${spec%%:*}; skip=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:cn_flat_kernel -s $skip -c 1 -f -o gpurun_out/${tag}_full_${wl} \
      python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_full_${wl}.log 2>&1
done
timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:cn_flat_kernel -s 145 -c 1 -f -o gpurun_out/${tag}_warm_c3 \
    python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_warm_c3.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
