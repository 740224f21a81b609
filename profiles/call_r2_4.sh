#!/bin/bash
# round 2, call 4: per-term gradient diagnostic (old fp32 kernels) + first run of the persistent kernels
mkdir -p gpurun_out
CRANK_B200_OPT_DISABLE=48 timeout 300 python profiles/diag_r2b.py fp32 16 500 1 2>&1 | grep -v Warn | tail -45 > gpurun_out/r2_diagb_fp32.txt; cat gpurun_out/r2_diagb_fp32.txt
CRANK_B200_OPT_DISABLE=48 timeout 300 python profiles/diag_r2b.py fp32 16 500 0 2>&1 | grep -v Warn | grep -A2 "^G:" > gpurun_out/r2_diagb_fp32_noragged.txt; cat gpurun_out/r2_diagb_fp32_noragged.txt
CRANK_B200_OPT_DISABLE=48 timeout 300 python profiles/diag_r2b.py fp32 2 96 1 2>&1 | grep -v Warn | grep -A2 "^G:" > gpurun_out/r2_diagb_fp32_toy.txt; cat gpurun_out/r2_diagb_fp32_toy.txt
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_pytest_tc.log; cat gpurun_out/r2_pytest_tc.log
timeout 200 python profiles/phase_probe.py tf32x3 2>&1 | grep -v diag | head -6 > gpurun_out/r2_phases_pt.txt; cat gpurun_out/r2_phases_pt.txt
for M in 0 32 48; do
  CRANK_B200_OPT_DISABLE=$M timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_mask$M.json 2> gpurun_out/r2_bench_mask$M.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_mask$M.json"))
    print("mask $M", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
except Exception as e:
    print("mask $M failed", e); print(open("gpurun_out/r2_bench_mask$M.err").read()[-1500:])
PY
done
