#!/bin/bash
# ncu evidence for bench.py (run under gpurun; writes into gpurun_out/)
#   1) launch list of one timed step region (device time per launch)
#   2) full capture of the dominant tensor-core kernels
set -x
PREC=${1:-tf32x3}
B=${2:-64}
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 2500 --csv \
    --log-file gpurun_out/launches_${PREC}.csv python bench.py --steps 1 --warmup 3 --precision $PREC --batch $B --no-cpu-baseline > gpurun_out/ncu_bench_${PREC}.log 2>&1
for K in k_resblock_fwd_tc k_conv_tc k_wgrad_tc; do
ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 2 \
    -o gpurun_out/prof_${K}_${PREC} -f python bench.py --steps 1 --warmup 3 --precision $PREC --batch $B --no-cpu-baseline > gpurun_out/ncu_full_${K}.log 2>&1
done
ls -la gpurun_out/
