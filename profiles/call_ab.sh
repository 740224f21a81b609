#!/bin/bash
# one gpurun call: GPU parity suite, A/B bench of the optional optimisations, phase probe, ncu launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for M in 0 15 14 7; do
  CRANK_B200_OPT_DISABLE=$M timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mask$M.json 2> gpurun_out/bench_mask$M.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_mask$M.json"))
    print("mask $M", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
except Exception as e:
    print("mask $M failed", e)
PY
done
timeout 200 python profiles/phase_probe.py tf32x3 > gpurun_out/phases_tf32x3.txt 2>&1; cat gpurun_out/phases_tf32x3.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 1000 --csv \
    --log-file gpurun_out/launches_tf32x3.csv python bench.py --steps 1 --warmup 3 --precision tf32x3 --no-cpu-baseline > gpurun_out/ncu_bench_tf32x3.log 2>&1
tail -2 gpurun_out/ncu_bench_tf32x3.log
