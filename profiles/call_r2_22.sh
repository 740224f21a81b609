#!/bin/bash
# round 2, call 22: wgrad: double-buffered G^T (k = 1 kernel) + TMA bulk-store epilogue; A/B (opt-disable 256 = direct stores)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_shapes.py tests/test_gpu_trainstep.py tests/test_gpu_graph.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2_pytest_22.log; cat gpurun_out/r2_pytest_22.log
for M in 0 256; do
CRANK_B200_OPT_DISABLE=$M timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_22_m$M.json 2> gpurun_out/r2_bench_22_m$M.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_22_m$M.json"))
    print("mask $M bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: (round(v["ms_per_step"], 2), round(v["avg_us"],1)) for k, v in d["kernels"].items()}, d["roofline"]["frac"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_22_m$M.err").read()[-1500:])
PY
done
