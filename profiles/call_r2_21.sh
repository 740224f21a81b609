#!/bin/bash
# round 2, call 21: wgrad split-once + integer RN split: whole GPU suite + bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest_21.log; cat gpurun_out/r2_pytest_21.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_21.json 2> gpurun_out/r2_bench_21.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_21.json"))
    print("bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", d["gpu_launches"], {k: (round(v["ms_per_step"], 2), round(v["avg_us"],1)) for k, v in d["kernels"].items()}, d["roofline"]["frac"], d["roofline"]["all_dense_kernels"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_21.err").read()[-1500:])
PY
