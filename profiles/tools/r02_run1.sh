#!/bin/bash
# Round 2, first GPU call (2 GPUs): every GPU test incl. the 2-GPU gather parity, the N=1 bench line, the N=2 bench in
# each gather mode.  Every step is bounded by its own timeout and writes under gpurun_out/.
mkdir -p gpurun_out
S=gpurun_out/r02_run1_summary.txt
: > $S
step() { local name=$1 limit=$2; shift 2; local t0=$(date +%s); timeout $limit "$@"; local rc=$?; echo "$name rc=$rc $(( $(date +%s) - t0 ))s" >> $S; }
nvidia-smi -L >> $S
step multi_tests 420 python -m pytest tests/test_gpu_multi.py tests/test_gpu_abi2.py -q -x > gpurun_out/r02_pytest_multi.log 2>&1
step bench_c2 300 bash -c 'python bench.py --steps 60 --warmup 5 > gpurun_out/r02_bench_c2_a.json 2> gpurun_out/r02_bench_c2_a.err'
port=29600
for mode in fused fused_mc fused_barrier nccl; do
  port=$((port+1))
  step bench_2gpu_$mode 300 bash -c "python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 60 --warmup 5 --gather $mode > gpurun_out/r02_bench_c2_2gpu_${mode}_a.json 2> gpurun_out/r02_bench_c2_2gpu_${mode}_a.err"
done
step gpu_tests 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_abi2.py > gpurun_out/r02_pytest_gpu.log 2>&1
cat $S
tail -5 gpurun_out/r02_pytest_multi.log gpurun_out/r02_pytest_gpu.log
for f in gpurun_out/r02_bench_c2_a.json gpurun_out/r02_bench_c2_2gpu_*_a.json; do echo == $f; head -c 3000 $f; echo; tail -5 ${f%.json}.err; done
