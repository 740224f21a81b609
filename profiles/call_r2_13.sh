#!/bin/bash
# round 2, call 13: new VQ kernels (parity + timing), conv-window test, whole GPU suite, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -m gpu -q -s -x -k "vq or ema or learnable" 2>&1 | grep -E "passed|failed|FAILED|Error|error|rel|assert|differ" | tail -30 > gpurun_out/r2_pytest_vq.log; cat gpurun_out/r2_pytest_vq.log
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest_all13.log; cat gpurun_out/r2_pytest_all13.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_13.json 2> gpurun_out/r2_bench_13.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_13.json"))
    print("bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", d["gpu_launches"], {k: (round(v["ms_per_step"], 2), round(v["avg_us"],1)) for k, v in d["kernels"].items()}, d["vq_argmin"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_13.err").read()[-1500:])
PY
