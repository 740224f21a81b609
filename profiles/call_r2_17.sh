#!/bin/bash
# round 2, call 17 (ONE GPU): ncu launch list of one train step + one --set full capture per kernel family (round-2 kernels)
mkdir -p gpurun_out
PREC=tf32x3
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 2900 -c 1000 --csv \
    --log-file gpurun_out/launches_r2_${PREC}.csv python bench.py --steps 1 --warmup 3 --precision $PREC --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/ncu_bench_r2.log 2>&1
tail -2 gpurun_out/ncu_bench_r2.log
for K in k_resblock_fwd_tc2 k_conv_tc k_wgrad_tc_raw k_vq_argmin_tf32; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K}\\b -s 40 -c 1 \
    -o gpurun_out/prof_r2_${K} -f python bench.py --steps 1 --warmup 3 --precision $PREC --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/ncu_full_r2_${K}.log 2>&1
tail -1 gpurun_out/ncu_full_r2_${K}.log
done
# the gate-backward instance of k_conv_tc (MODE 1)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc<.*1, .*1>" -s 20 -c 1 \
    -o gpurun_out/prof_r2_k_conv_tc_gate -f python bench.py --steps 1 --warmup 3 --precision $PREC --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/ncu_full_r2_gate.log 2>&1
tail -1 gpurun_out/ncu_full_r2_gate.log
ls -la gpurun_out/ | grep r2_ | tail -12
