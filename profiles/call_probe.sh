#!/bin/bash
# one gpurun call: GPU parity suite + phase probe + short A/B bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -15
timeout 200 python profiles/phase_probe.py tf32x3 > gpurun_out/phases_tf32x3.txt 2>&1; cat gpurun_out/phases_tf32x3.txt
for M in ${MASKS:-0 16}; do
  CRANK_B200_OPT_DISABLE=$M timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mask$M.json 2> gpurun_out/bench_mask$M.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_mask$M.json"))
    print("mask $M", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
except Exception as e:
    print("mask $M failed", e); print(open("gpurun_out/bench_mask$M.err").read()[-2000:])
PY
done
