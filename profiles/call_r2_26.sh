#!/bin/bash
# round 2, call 26: 16 worker warps in k_wgrad_tc_raw: parity + A/B (opt-disable 1024 = 8 warps) + phase stamps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_shapes.py tests/test_gpu_trainstep.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_pytest_26.log; cat gpurun_out/r2_pytest_26.log
for M in 0 2048; do
CRANK_B200_OPT_DISABLE=$M timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_26_m$M.json 2> gpurun_out/r2_bench_26_m$M.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_26_m$M.json"))
    print("mask $M bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: (round(v["ms_per_step"], 2), round(v["avg_us"],1)) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_26_m$M.err").read()[-1500:])
PY
done
timeout 200 python profiles/phase_probe.py tf32x3 2>&1 | grep -v Warn | grep -B2 "wgrad" | cut -c1-600
