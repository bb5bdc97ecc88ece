#!/bin/bash
# Round 2 (second session) diagnostic: ncu --set full of the step kernel WITHOUT cache flushes between replay passes
# (--cache-control none: instruction caches warm, as in a graph-replayed rollout) at c2 and c3 -- how much of the
# stalled_no_instruction figure of the committed (cold) captures is ncu's own flush?
mkdir -p gpurun_out
for spec in c2:529 c3:145; do
  wl=${spec%%:*}; skip=${spec##*:}
  timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:cn_flat_kernel -s $skip -c 1 -f \
      -o gpurun_out/r02b_warm_${wl} python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras \
      > gpurun_out/r02b_warm_${wl}.log 2>&1
  echo "$wl rc=$?"
done
ls -la gpurun_out/*.ncu-rep
