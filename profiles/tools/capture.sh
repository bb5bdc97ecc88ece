#!/bin/bash
# ncu evidence for one round: launch list of a bench run + one full capture of the step kernel.
# usage (GPU box): bash profiles/tools/capture.sh <tag> <workload>
tag=$1; wl=${2:-c2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_${wl}.csv \
    python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_${wl}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cn_flat_kernel -s 30 -c 1 -o gpurun_out/${tag}_full_${wl} \
    python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_full_${wl}.log 2>&1
ls -la gpurun_out | tail -5
