#!/bin/bash
# one gpurun call (ONE GPU): phase probe + ncu launch list of one train step + one --set full capture per kernel family
set -x
mkdir -p gpurun_out
PREC=${1:-tf32x3}
timeout 200 python profiles/phase_probe.py $PREC > gpurun_out/phases_$PREC.txt 2>&1; cat gpurun_out/phases_$PREC.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1000 --csv \
    --log-file gpurun_out/launches_${PREC}.csv python bench.py --steps 1 --warmup 3 --precision $PREC --no-cpu-baseline > gpurun_out/ncu_bench_${PREC}.log 2>&1
for K in k_resblock_fwd_tc k_conv_tc k_wgrad_tc_raw k_wgrad_tc k_vq_argmin; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K}\\b -s 40 -c 1 \
    -o gpurun_out/prof_${K}_${PREC} -f python bench.py --steps 1 --warmup 3 --precision $PREC --no-cpu-baseline > gpurun_out/ncu_full_${K}.log 2>&1
tail -2 gpurun_out/ncu_full_${K}.log
done
ls -la gpurun_out/ | head -40
