#!/bin/bash
# round 2, call 14: K-phased 2-CTA/SM k_conv_tc (dgrad / gate backward / plain conv), fused EMA diagnosis, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -x -k "fused_ema" 2>&1 | grep -E "passed|failed|FAILED|Error|error|differs|assert" | tail -12
timeout 1800 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_shapes.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_pytest_14.log; cat gpurun_out/r2_pytest_14.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_14.json 2> gpurun_out/r2_bench_14.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_14.json"))
    print("bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", d["gpu_launches"], {k: (round(v["ms_per_step"], 2), round(v["avg_us"],1)) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_14.err").read()[-1500:])
PY
