#!/bin/bash
# round 2, call 28 (ONE GPU): final evidence with the final kernels: call 24 (suite, smoke, bench lines, phases) + ncu launch list and
# --set full captures exported as CSV (call 17 / 18)
bash profiles/call_r2_24.sh
mkdir -p gpurun_out /tmp/ncu
PREC=tf32x3
B="python bench.py --steps 1 --warmup 3 --precision $PREC --no-cpu-baseline --no-eager-gpu-baseline"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 2900 -c 1000 --csv \
    --log-file gpurun_out/launches_r2_${PREC}.csv $B > gpurun_out/ncu_bench_r2.log 2>&1
cap() {   # name, kernel regex (demangled name incl. template arguments), skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$2" -s $3 -c 1 -o /tmp/ncu/$1 -f $B > gpurun_out/ncu_full_r2_$1.log 2>&1
  tail -1 gpurun_out/ncu_full_r2_$1.log
  ncu -i /tmp/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu_r2_$1_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$1.ncu-rep --page source --csv > gpurun_out/ncu_r2_$1_source.csv 2>/dev/null
  rm -f /tmp/ncu/$1.ncu-rep
}
cap k_resblock_fwd_tc2 'k_resblock_fwd_tc2' 40
cap k_conv_tc_dgrad 'k_conv_tc<.bool.1, .int.2' 40
cap k_conv_tc_gate 'k_conv_tc<.bool.1, .int.1' 20
cap k_wgrad_tc_raw 'k_wgrad_tc_raw' 20
cap k_wgrad_tc 'k_wgrad_tc<' 20
cap k_vq_argmin_tf32 'k_vq_argmin_tf32' 4
du -sh gpurun_out
