#!/bin/bash
# round 2, call 18: VQ scan rewrite (parity + timing), whole GPU suite, bench, dgrad / gate ncu captures
mkdir -p gpurun_out /tmp/ncu
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest_18.log; cat gpurun_out/r2_pytest_18.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_18.json 2> gpurun_out/r2_bench_18.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_18.json"))
    print("bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", d["gpu_launches"], {k: (round(v["ms_per_step"], 2), round(v["avg_us"],1)) for k, v in d["kernels"].items()}, d["vq_argmin"]["frac"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_18.err").read()[-1500:])
PY
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline"
cap() {   # name, kernel regex (demangled name incl. template arguments), skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$2" -s $3 -c 1 -o /tmp/ncu/$1 -f $B > gpurun_out/ncu_full_r2_$1.log 2>&1
  tail -1 gpurun_out/ncu_full_r2_$1.log
  ncu -i /tmp/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu_r2_$1_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$1.ncu-rep --page source --csv > gpurun_out/ncu_r2_$1_source.csv 2>/dev/null
  rm -f /tmp/ncu/$1.ncu-rep
}
cap k_conv_tc_dgrad 'k_conv_tc<.bool.1, .int.2' 40
cap k_conv_tc_gate 'k_conv_tc<.bool.1, .int.1' 20
cap k_vq_argmin_tf32 'k_vq_argmin_tf32' 4
ls -la gpurun_out | tail -8
