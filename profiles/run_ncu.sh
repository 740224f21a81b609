#!/bin/bash
# ncu evidence for bench.py (run under gpurun, ONE GPU; writes into gpurun_out/)
#   1) launch list covering one full train step after the warm-up steps (device time per launch;
#      cold-cache + serialised: compare SHARES, not absolutes)
#   2) one full capture (--set full, source) of each tensor-core kernel family
set -x
PREC=${1:-tf32x3}
B=${2:-64}
ncu --metrics gpu__time_duration.sum --clock-control none -s 3700 -c 1300 --csv \
    --log-file gpurun_out/launches_${PREC}.csv python bench.py --steps 1 --warmup 3 --precision $PREC --batch $B --no-cpu-baseline > gpurun_out/ncu_bench_${PREC}.log 2>&1
for K in k_resblock_fwd_tc k_conv_tc k_wgrad_tc k_vq_argmin; do
ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 1 \
    -o gpurun_out/prof_${K}_${PREC} -f python bench.py --steps 1 --warmup 3 --precision $PREC --batch $B --no-cpu-baseline > gpurun_out/ncu_full_${K}.log 2>&1
done
ls -la gpurun_out/ | head -30
